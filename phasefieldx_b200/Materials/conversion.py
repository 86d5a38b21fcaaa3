"""Elastic-constant conversions, API of reference `src/phasefieldx/Materials/conversion.py:10-144`
(same names, argument meaning and ValueError behaviour)."""


def _check_nu(nu):
    if not (-1 < nu < 0.5):
        raise ValueError("Poisson's ratio (nu) must be within the range -1 < nu < 0.5")


def get_lambda_lame(E, nu):
    _check_nu(nu)
    return E * nu / ((1.0 - 2.0 * nu) * (1.0 + nu))


def get_mu_lame(E, nu):
    _check_nu(nu)
    return E / (2.0 * (1.0 + nu))


def get_bulk_modulus(E, nu):
    _check_nu(nu)
    return E / (3.0 * (1.0 - 2.0 * nu))


def get_youngs_modulus(lambda_, mu):
    if mu <= 0:
        raise ValueError("Lame's second parameter (mu) must be positive")
    return mu * (3 * lambda_ + 2 * mu) / (lambda_ + mu)


def get_poissons_ratio(lambda_, mu):
    if lambda_ + 2 * mu == 0:
        raise ValueError("Invalid combination of Lame's parameters (lambda and mu)")
    return lambda_ / (2 * (lambda_ + mu))


def get_bulk_modulus_lame(lambda_, mu):
    if lambda_ + 2 * mu == 0:
        raise ValueError("Invalid combination of Lame's parameters (lambda and mu)")
    return lambda_ + (2 / 3) * mu
