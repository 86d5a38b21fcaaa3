"""Shared host-side scaffolding of the single-field drop-in solvers (results folder, logger, VTU)."""
import os
import time

from phasefieldx_b200.files import prepare_simulation
from phasefieldx_b200.io_vtk import VTKFile
from phasefieldx_b200.io_xdmf import XDMFFile
from phasefieldx_b200.Logger.library_versions import (log_end_analysis, log_library_versions, log_model_information,
                                                      log_system_info, set_logger)


class Run:
    """Set-up shared by reference Elasticity/solver/solver.py:78-107 and Phase_Field/solver/solver.py:84-101."""

    def __init__(self, Data, msh, path, options, vtu_format="ascii"):
        self.path = os.getcwd() if path is None else path
        self.folder = Data.results_folder_name
        prepare_simulation(self.path, self.folder)
        self.logger = set_logger(self.folder)
        log_system_info(self.logger)
        log_library_versions(self.logger)
        Data.save_log_info(self.logger)
        Data.save_parameters_to_csv(os.path.join(self.folder, "parameters.input"))
        log_model_information(msh, self.logger)
        self.logger.info(" SNES Settings:")
        for key, value in options.items():
            self.logger.info(f"   {key}: {value}")
        self.start = time.perf_counter()
        self.logger.info(f" start time: {self.start}")
        self.vtk = None
        if Data.save_solution_vtu:
            self.vtk = VTKFile(None, os.path.join(self.folder, "paraview-solutions_vtu", "phasefieldx.pvd"), "w", fmt=vtu_format)
        self.xdmf = {}
        self._xdmf_on = bool(getattr(Data, "save_solution_xdmf", False))
        self._msh = msh
        self.logger.info(" S t a r t i n g    A n a l y s i s ")
        self.logger.info(" ---------------------------------- ")
        self.logger.info(" ---------------------------------- ")

    def step_header(self, t, dt, step):
        self.logger.info(f"\n\nSolution at (pseudo) time = {t}, dt = {dt}, Step = {step} ")
        self.logger.info("===========================================================================")

    def write_xdmf(self, fields, step):
        """`<name>.xdmf` per field in paraview-solutions_xdmf/ (reference xdmf_phi / xdmf_u), when Data.save_solution_xdmf."""
        if not self._xdmf_on:
            return
        for name, values in fields.items():
            if name not in self.xdmf:
                self.xdmf[name] = XDMFFile(None, os.path.join(self.folder, "paraview-solutions_xdmf", name + ".xdmf"), "w")
                self.xdmf[name].write_mesh(self._msh)
            self.xdmf[name].write_function(name, values, step)

    def finish(self):
        if self.vtk is not None:
            self.vtk.close()
        for xf in self.xdmf.values():
            xf.close()
        log_end_analysis(self.logger, time.perf_counter() - self.start)
