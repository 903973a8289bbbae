#!/usr/bin/env python
"""Convert a Mann turbulence box written by hipersim / dynamiks (``MannTurbulenceField.to_netcdf``: NetCDF-4, i.e.
HDF5) into the ``.npz`` layout ``windgym_b200.mann.MannBox.from_file`` reads (arrays ``uvw`` [3, Nx, Ny, Nz] and
``dxyz`` [3]).  Run it where the box was made -- it needs ONE of netCDF4, h5py or xarray, none of which the GPU image
has (reference call sites: Wind_Farm_Env.py:611-618 ``from_netcdf``, FarmEval.py:86-90 ``update_tf``).

    python scripts/convert_mann_netcdf.py TF_files/*.nc          # writes TF_files/*.npz next to the inputs
"""
import sys

import numpy as np


def read_box(path):
    try:
        import netCDF4
        with netCDF4.Dataset(path) as nc:
            return np.asarray(nc["uvw"][:]), [np.asarray(nc[a][:]) for a in ("x", "y", "z")]
    except ImportError:
        pass
    try:
        import h5py
        with h5py.File(path, "r") as f:
            return np.asarray(f["uvw"]), [np.asarray(f[a]) for a in ("x", "y", "z")]
    except ImportError:
        pass
    import xarray as xr
    ds = xr.load_dataset(path)
    da = ds["uvw"] if "uvw" in ds else ds.to_array().squeeze()
    return np.asarray(da), [np.asarray(ds[a]) for a in ("x", "y", "z")]


def main(paths):
    for p in paths:
        uvw, axes = read_box(p)
        dxyz = np.array([float(a[1] - a[0]) for a in axes])
        out = p.rsplit(".", 1)[0] + ".npz"
        np.savez(out, uvw=uvw.astype(np.float32), dxyz=dxyz)
        print(f"{p} -> {out}: uvw {uvw.shape}, dxyz {dxyz.tolist()}")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    main(sys.argv[1:])
