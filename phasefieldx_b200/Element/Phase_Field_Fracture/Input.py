"""Simulation parameters, API of reference `src/phasefieldx/Element/Phase_Field_Fracture/Input.py:10-181`
(same keyword names and defaults, `parameters.input` TSV format, log lines)."""
from phasefieldx_b200.Materials.conversion import get_lambda_lame, get_mu_lame

_ORDER = ("E", "nu", "Gc", "l", "k", "lambda", "mu", "degradation", "split_energy", "degradation_function",
          "irreversibility", "fatigue", "fatigue_degradation_function", "fatigue_val", "save_solution_xdmf",
          "save_solution_vtu", "results_folder_name")


class Input:
    def __init__(self, E=210.0, nu=0.3, Gc=0.0027, l=0.015, degradation="isotropic", split_energy="no",
                 degradation_function="quadratic", irreversibility="no", fatigue=False,
                 fatigue_degradation_function="asymptotic", fatigue_val=0.05625, k=0, save_solution_xdmf=False,
                 save_solution_vtu=True, results_folder_name="results"):
        self.E, self.nu, self.Gc, self.l, self.k = E, nu, Gc, l, k
        self.lambda_ = get_lambda_lame(E, nu)
        self.mu = get_mu_lame(E, nu)
        self.degradation = degradation
        self.split_energy = split_energy
        self.degradation_function = degradation_function
        self.irreversibility = irreversibility
        self.fatigue = fatigue
        self.fatigue_degradation_function = fatigue_degradation_function
        self.fatigue_val = fatigue_val
        self.save_solution_xdmf = save_solution_xdmf
        self.save_solution_vtu = save_solution_vtu
        self.results_folder_name = results_folder_name

    def _as_dict(self):
        d = {name: getattr(self, name, None) for name in _ORDER}
        d["lambda"] = self.lambda_
        return d

    def _lines(self):
        return ["Material parameters:", f"  E: {self.E}", f"  nu: {self.nu}", f"  Gc: {self.Gc}", f"  l: {self.l}",
                f"  k: {self.k}", f"  lambda: {self.lambda_}", f"  mu: {self.mu}", "Phase-field fracture model:",
                f"  degradation: {self.degradation}", f"  split_energy: {self.split_energy}",
                f"  degradation_function: {self.degradation_function}", f"  irreversibility: {self.irreversibility}",
                f"  fatigue degradation function: {self.fatigue_degradation_function}",
                f"  fatigue_val: {self.fatigue_val}"]

    def save_log_info(self, logger):
        for line in self._lines():
            logger.info(line)

    def save_parameters_to_csv(self, filename="parameters.input"):
        with open(filename, "w") as fh:
            for key, value in self._as_dict().items():
                fh.write(f"{key}\t{value}\n")

    def __str__(self):
        return "\n".join(self._lines())
