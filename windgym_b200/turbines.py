"""Tabular turbine handed to the env as ``turbine=`` (reference: py_wake ``V80``, ``tests/test_basics.py:4,:17``).

py_wake is not a dependency of this package; any object with ``diameter()``, ``hub_height()``, ``power(ws)`` and
the three tables works (a real py_wake turbine can be wrapped with ``TabularTurbine.from_pywake``).
"""
import numpy as np


class TabularTurbine:
    def __init__(self, name, diameter, hub_height, ws, power_w, ct):
        self.name = name
        self._d, self._h = float(diameter), float(hub_height)
        self.ws_table = np.asarray(ws, dtype=np.float64)
        self.power_table_w = np.asarray(power_w, dtype=np.float64)
        self.ct_table = np.asarray(ct, dtype=np.float64)

    def diameter(self):
        return self._d

    def hub_height(self):
        return self._h

    def power(self, ws, yaw=0.0):
        """P(ws cos yaw) [W] -- py_wake SimpleYawModel semantics."""
        return np.interp(np.asarray(ws, dtype=np.float64) * np.cos(np.deg2rad(yaw)), self.ws_table, self.power_table_w)

    def ct(self, ws, yaw=0.0):
        co = np.cos(np.deg2rad(yaw))
        return np.interp(np.asarray(ws, dtype=np.float64) * co, self.ws_table, self.ct_table) * co ** 2

    @classmethod
    def from_pywake(cls, wt, ws=np.arange(3.0, 26.0, 1.0)):
        return cls(getattr(wt, "name", lambda: "wt")(), wt.diameter(), wt.hub_height(), ws, wt.power(ws), wt.ct(ws))


def V80():
    """Vestas V80-2.0MW, D = 80 m, hub 70 m (py_wake ``examples/data/hornsrev1.py`` tables)."""
    p_kw = [0.0, 66.6, 154.0, 282.0, 460.0, 696.0, 996.0, 1341.0, 1661.0, 1866.0, 1958.0, 1988.0, 1997.0, 1999.0] + [2000.0] * 9
    ct = [0.0, 0.818, 0.806, 0.804, 0.805, 0.806, 0.807, 0.793, 0.739, 0.709, 0.409, 0.314, 0.249, 0.202,
          0.167, 0.140, 0.119, 0.102, 0.088, 0.077, 0.067, 0.060, 0.053]
    return TabularTurbine("V80", 80.0, 70.0, np.arange(3.0, 26.0, 1.0), np.array(p_kw) * 1000.0, ct)
