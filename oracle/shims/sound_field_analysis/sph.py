"""Shim for sound_field_analysis.sph (oracle/test infrastructure only)."""
from scipy.special import spherical_jn, spherical_yn


def sphankel2(n, kr):
    """Spherical Hankel function of the second kind, h_n^(2)(kr) = j_n(kr) - i y_n(kr)."""
    return spherical_jn(n, kr) - 1j * spherical_yn(n, kr)
