"""CPU: the Mann turbulence box -- oracle statistics, the torch generator against the oracle on the
same noise (run on CPU tensors here; the GPU run is in test_gpu_turbulence.py), box layouts and sampling."""
import numpy as np
import torch

from oracle import mann_numpy as mn
from windgym_b200.mann import MannBox


def test_oracle_box_has_mann_statistics():
    uvw = mn.mann_box(0.1, 33.6, 3.9, (128, 32, 32), (4.0, 4.0, 4.0), seed=5)
    assert uvw.shape == (3, 128, 32, 32) and np.abs(uvw.mean(axis=(1, 2, 3))).max() < 1e-12
    su, sv, sw = uvw.std(axis=(1, 2, 3))
    assert su > sv > sw > 0.3 * su                 # sheared tensor (Gamma = 3.9): sigma_u > sigma_v > sigma_w
    assert np.mean(uvw[0] * uvw[2]) < 0            # negative uw co-variance (momentum flux towards the ground)
    iso = mn.mann_box(0.1, 33.6, 0.0, (64, 32, 32), (4.0, 4.0, 4.0), seed=5)
    s = iso.std(axis=(1, 2, 3))
    assert abs(s[1] / s[2] - 1) < 0.1 and abs(np.mean(iso[0] * iso[2])) < 0.05 * s[0] * s[2]   # Gamma = 0: isotropic
    # energy scales with alphaepsilon (amplitude with its square root)
    twice = mn.mann_box(0.4, 33.6, 3.9, (128, 32, 32), (4.0, 4.0, 4.0), seed=5)
    assert np.allclose(twice, 2.0 * uvw)


def test_torch_generator_matches_oracle_on_the_same_noise():
    N, d = (32, 16, 8), (3.0, 4.0, 5.0)
    noise = mn.box_noise(N, 7)
    ref = mn.mann_box(0.1, 33.6, 3.9, N, d, noise=noise)
    box = MannBox.generate(0.1, 33.6, 3.9, N, d, device="cpu", noise=noise, lowpass_width=12.0, slab=8)
    got = box.raw[..., :3].permute(3, 0, 1, 2).numpy()
    assert np.abs(got - ref).max() < 1e-5 * ref.std()
    assert np.all(box.raw[..., 3].numpy() == 0) and box.raw.shape == N + (4,) and box.lp.shape == N + (2,)
    f = mn.MannTurbulenceField(ref, d, lowpass_width=12.0)
    assert np.abs(box.lp.permute(3, 0, 1, 2).numpy() - f.uvw_lp[1:]).max() < 1e-5 * ref.std()
    assert abs(box.std_u - ref[0].std()) < 1e-6
    f.scale_TI(0.1, 9.0)
    assert np.isclose(box.scale_for(0.1, 9.0), f.scale) and np.isclose(np.std(f.uvw[0]) * f.scale, 0.9)
    # a seeded device draw is reproducible and differs between seeds
    a = MannBox.generate(Nxyz=(16, 8, 8), dxyz=(3.0, 3.0, 3.0), seed=3, device="cpu")
    b = MannBox.generate(Nxyz=(16, 8, 8), dxyz=(3.0, 3.0, 3.0), seed=3, device="cpu")
    c = MannBox.generate(Nxyz=(16, 8, 8), dxyz=(3.0, 3.0, 3.0), seed=4, device="cpu")
    assert torch.equal(a.raw, b.raw) and not torch.equal(a.raw, c.raw)


def test_field_sampling_is_periodic_trilinear_and_taylor_shifted(tmp_path):
    uvw = mn.mann_box(0.1, 33.6, 3.9, (16, 8, 8), (3.0, 3.0, 3.0), seed=2)
    f = mn.MannTurbulenceField(uvw, (3.0, 3.0, 3.0), lowpass_width=9.0, offset=(1.5, 0.0, 3.0))
    f.scale_TI(0.08, 10.0)
    # grid nodes reproduce the box, a mid point is the mean of its two neighbours, the box is periodic
    assert np.allclose(f.sample(-1.5, 0.0, -3.0, 0.0, 10.0)[:, None], f.scale * uvw[:, :1, 0, 0])
    mid = f.sample(0.0, 0.0, -3.0, 0.0, 10.0)
    assert np.allclose(mid, f.scale * 0.5 * (uvw[:, 0, 0, 0] + uvw[:, 1, 0, 0]))
    assert np.allclose(f.sample(7.0, 5.0, 2.0, 0.0, 10.0), f.sample(7.0 + 48.0, 5.0 - 24.0, 2.0 + 24.0, 0.0, 10.0))
    # frozen turbulence: what is at x now is at x + U dt a moment later
    assert np.allclose(f.sample(4.0, 1.0, 2.0, 3.0, 10.0), f.sample(4.0 + 20.0, 1.0, 2.0, 5.0, 10.0))
    assert f.sample_lp(np.zeros(4), np.zeros(4), np.zeros(4), 0.0, 10.0).shape == (2, 4)
    # file round trip of the device-side loader
    np.savez(tmp_path / "box.npz", uvw=uvw.astype(np.float32), dxyz=np.array([3.0, 3.0, 3.0]))
    box = MannBox.from_file(str(tmp_path / "box.npz"), device="cpu", lowpass_width=9.0)
    assert box.Nxyz == (16, 8, 8) and np.allclose(box.raw[..., :3].permute(3, 0, 1, 2).numpy(), uvw, atol=1e-6)
