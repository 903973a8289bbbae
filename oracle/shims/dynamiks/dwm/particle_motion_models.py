from oracle.dwm_numpy import HillVortexParticleMotion  # noqa: F401
