"""Simulation parameters, API of reference `src/phasefieldx/Element/Phase_Field/Input.py:33-90`."""


class Input:
    def __init__(self, l=1.0, save_solution_xdmf=False, save_solution_vtu=True, results_folder_name="results"):
        self.l = l
        self.save_solution_xdmf = save_solution_xdmf
        self.save_solution_vtu = save_solution_vtu
        self.results_folder_name = results_folder_name

    def save_log_info(self, logger):
        logger.info("Parameters:")
        logger.info(f"  l: {self.l}")

    def save_parameters_to_csv(self, filename="parameters.input"):
        params = {"l": self.l, "save_solution_xdmf": self.save_solution_xdmf,
                  "save_solution_vtu": self.save_solution_vtu, "results_folder_name": self.results_folder_name}
        with open(filename, "w") as fh:
            for key, value in params.items():
                fh.write(f"{key}\t{value}\n")

    def __str__(self):
        return "\n".join(["Parameters:", f"  l: {self.l}"])
