"""TEST INFRASTRUCTURE -- restatement of py_wake's ``V80`` (``py_wake/examples/data/hornsrev1.py``).

py_wake is an un-vendored, unpinned dependency of the reference (pulled by ``dynamiks``,
``/root/reference/pyproject.toml:25``; used at ``tests/test_basics.py:4``, ``Wind_Farm_Env.py:112,:244,:700``).
The tables are the published Vestas V80-2.0MW curves as shipped by py_wake.
Interpolation is linear (``np.interp`` semantics: clamped to the end values outside 3..25 m/s).
"""
import numpy as np

V80_WS = np.arange(3.0, 26.0, 1.0)  # 23 knots, 3..25 m/s
V80_POWER_KW = np.array(
    [0.0, 66.6, 154.0, 282.0, 460.0, 696.0, 996.0, 1341.0, 1661.0, 1866.0, 1958.0, 1988.0, 1997.0, 1999.0]
    + [2000.0] * 9
)
V80_CT = np.array(
    [0.0, 0.818, 0.806, 0.804, 0.805, 0.806, 0.807, 0.793, 0.739, 0.709, 0.409, 0.314, 0.249, 0.202,
     0.167, 0.140, 0.119, 0.102, 0.088, 0.077, 0.067, 0.060, 0.053]
)
assert V80_WS.size == V80_POWER_KW.size == V80_CT.size == 23


class V80:
    """Duck-type of the py_wake turbine object the env receives as ``turbine=`` (``Wind_Farm_Env.py:52``)."""

    name = "V80"

    def diameter(self):
        return 80.0

    def hub_height(self):
        return 70.0

    def power(self, ws, yaw=0.0):
        """Electrical power [W]; py_wake ``SimpleYawModel``: P(ws*cos(yaw))."""
        ws = np.asarray(ws, dtype=np.float64) * np.cos(np.deg2rad(yaw))
        return np.interp(ws, V80_WS, V80_POWER_KW * 1000.0)

    def ct(self, ws, yaw=0.0):
        """Thrust coefficient; ``SimpleYawModel``: CT(ws*cos(yaw)) * cos(yaw)**2."""
        co = np.cos(np.deg2rad(yaw))
        ws = np.asarray(ws, dtype=np.float64) * co
        return np.interp(ws, V80_WS, V80_CT) * co**2

    # tables for the device side (same numbers, one source of truth)
    ws_table = V80_WS
    power_table_w = V80_POWER_KW * 1000.0
    ct_table = V80_CT
