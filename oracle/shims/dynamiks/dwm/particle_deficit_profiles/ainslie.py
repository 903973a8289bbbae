from oracle.dwm_numpy import jDWMAinslieGenerator  # noqa: F401
