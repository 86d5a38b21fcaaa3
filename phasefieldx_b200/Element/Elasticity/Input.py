"""Simulation parameters, API of reference `src/phasefieldx/Element/Elasticity/Input.py:46-111`
(same keyword names and defaults, `parameters.input` TSV format, log lines)."""
from phasefieldx_b200.Materials.conversion import get_lambda_lame, get_mu_lame


class Input:
    def __init__(self, E=210.0, nu=0.3, save_solution_xdmf=False, save_solution_vtu=True, results_folder_name="results"):
        self.E, self.nu = E, nu
        self.lambda_ = get_lambda_lame(E, nu)
        self.mu = get_mu_lame(E, nu)
        self.save_solution_xdmf = save_solution_xdmf
        self.save_solution_vtu = save_solution_vtu
        self.results_folder_name = results_folder_name

    def _lines(self):
        return ["Material parameters:", f"  E: {self.E}", f"  nu: {self.nu}", f"  lambda: {self.lambda_}", f"  mu: {self.mu}"]

    def save_log_info(self, logger):
        for line in self._lines():
            logger.info(line)

    def save_parameters_to_csv(self, filename="parameters.input"):
        params = {"E": self.E, "nu": self.nu, "lambda": self.lambda_, "mu": self.mu,
                  "save_solution_xdmf": self.save_solution_xdmf, "save_solution_vtu": self.save_solution_vtu,
                  "results_folder_name": self.results_folder_name}
        with open(filename, "w") as fh:
            for key, value in params.items():
                fh.write(f"{key}\t{value}\n")

    def __str__(self):
        return "\n".join(self._lines())
