"""CPU: host logic of the f-2 / f-3 rows -- agents, vector-env adapters, episode statistics, batched evaluation
assembly -- on a fake ``VecWindFarmEnv`` (tests/fake_vec_env.py).  The CUDA-backed versions run in test_gpu_adapters."""
import numpy as np
import pytest
import torch

from tests.fake_vec_env import FakeVecEnv
from windgym_b200.agents import BaseAgent, ConstantAgent, GreedyAgent, RandomAgent, SB3MlpPolicy, batch_actions
from windgym_b200.evaluate import AgentEval, EvalDataset, condition_grid, eval_batched
from windgym_b200.vector import GymVectorEnv, RecordEpisodeVals, SB3VecEnv


def test_scale_yaw_matches_reference_known_answer():
    # ConstantAgent([-10, 20, 0, 0]) on yaw range [-45, 45] -> action [-0.2222, 0.4444, 0, 0]
    a, st = ConstantAgent([-10, 20, 0, 0]).predict()
    assert st is None
    assert np.allclose(a, [-0.22222222, 0.44444444, 0.0, 0.0])
    assert BaseAgent(30, -30).scale_yaw(np.array([30.0, -30.0, 0.0])).tolist() == [1.0, -1.0, 0.0]


def test_batched_agents():
    env = FakeVecEnv(4, n_turb=3)
    obs, _ = env.reset()
    a = batch_actions(ConstantAgent([-10, 20, 0]), obs, env)
    assert a.shape == (4, 3) and torch.allclose(a[2], torch.tensor([-0.22222222, 0.44444444, 0.0]))
    r = batch_actions(RandomAgent(env=env, seed=3), obs, env)
    assert r.shape == (4, 3) and float(r.abs().max()) <= 1.0
    env.state["yaw"][:, 0] = torch.tensor([[5.0, -0.5, 0.0]] * 4)
    g = GreedyAgent(type="global", env=env)
    goal = (g.predict_batch(obs, env) + 1) / 2 * 90 - 45
    assert torch.allclose(goal[0], torch.tensor([4.0, 0.0, 0.0]), atol=1e-5)

    class PerEnv:  # SB3-style predict(obs) agent: called once per env
        def predict(self, o, deterministic=False):
            return np.full(3, o[0], dtype=np.float32), None
    p = batch_actions(PerEnv(), obs, env)
    assert torch.allclose(p[:, 0], obs[:, 0])


def test_sb3_mlp_policy_from_state_dict():
    g = torch.Generator().manual_seed(0)
    sd = {"log_std": torch.zeros(4),
          "mlp_extractor.policy_net.0.weight": torch.randn(64, 8, generator=g), "mlp_extractor.policy_net.0.bias": torch.randn(64, generator=g),
          "mlp_extractor.policy_net.2.weight": torch.randn(64, 64, generator=g) * 0.1, "mlp_extractor.policy_net.2.bias": torch.zeros(64),
          "mlp_extractor.value_net.0.weight": torch.randn(64, 8, generator=g), "mlp_extractor.value_net.0.bias": torch.zeros(64),
          "action_net.weight": torch.randn(4, 64, generator=g) * 0.1, "action_net.bias": torch.zeros(4),
          "value_net.weight": torch.randn(1, 64, generator=g), "value_net.bias": torch.zeros(1)}
    pol = SB3MlpPolicy.from_state_dict(sd)
    obs = torch.randn(5, 8, generator=g)
    h = torch.tanh(obs @ sd["mlp_extractor.policy_net.0.weight"].T + sd["mlp_extractor.policy_net.0.bias"])
    h = torch.tanh(h @ sd["mlp_extractor.policy_net.2.weight"].T)
    ref = (h @ sd["action_net.weight"].T).clamp(-1, 1)
    assert torch.allclose(pol.predict_batch(obs), ref, atol=1e-6)
    a, _ = pol.predict(obs[0].numpy())
    assert a.shape == (4,) and np.allclose(a, ref[0].numpy(), atol=1e-6)
    with pytest.raises(ValueError):
        SB3MlpPolicy.from_state_dict({"action_net.weight": torch.zeros(4, 8)})


def test_gym_vector_env_same_step_autoreset():
    fake = FakeVecEnv(3, horizon=4)
    env = GymVectorEnv(venv=fake)
    assert env.num_envs == 3 and env.single_action_space.shape == (3,) and env.single_observation_space.shape == (6,)
    with pytest.raises(RuntimeError):
        env.step(np.zeros((3, 3), dtype=np.float32))
    obs, infos = env.reset(seed=1)
    assert obs.shape == (3, 6) and obs.dtype == np.float32 and infos["Power agent"].shape == (3,)
    fake.state["timestep"][1] = 2          # env 1 is two steps ahead: it finishes at the second step
    seen = []
    for k in range(4):
        obs, r, term, trunc, infos = env.step(np.ones((3, 3), dtype=np.float32))
        assert r.dtype == np.float64 and trunc.dtype == bool and not term.any()
        seen.append(trunc.copy())
        if trunc.any():
            assert np.array_equal(infos["_final_observation"], trunc)
            fo = infos["final_observation"]
            # finished envs were reset in the same step: their new obs has yaw 0, the final obs the stepped yaw
            assert np.all(obs[trunc][:, 3:] == 0.0) and np.all(fo[trunc][:, 3:] > 0.0)
        else:
            assert "final_observation" not in infos
    assert [s.tolist() for s in seen] == [[False, False, False], [False, True, False], [False, False, False],
                                         [True, False, True]]
    assert fake.n_resets.tolist() == [2, 2, 2]


def test_sb3_vec_env_protocol():
    fake = FakeVecEnv(2, horizon=2)
    env = SB3VecEnv(venv=fake)
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (2, 6)
    env.step_async(np.zeros((2, 3), dtype=np.float32))
    obs, rew, dones, infos = env.step_wait()
    assert not dones.any() and len(infos) == 2 and "terminal_observation" not in infos[0]
    assert infos[1]["Power agent"] == pytest.approx(float(fake.state["power"][1, 0].sum()))
    obs, rew, dones, infos = env.step(np.zeros((2, 3), dtype=np.float32))
    assert dones.all() and infos[0]["TimeLimit.truncated"] and infos[0]["terminal_observation"].shape == (6,)
    assert env.get_attr("n_turb") == [3, 3] and env.env_is_wrapped(object) == [False, False]
    assert env.seed(7) == [7, 7] and fake.seed == 7
    env.close()
    assert fake.closed


def test_record_episode_vals_mean_power():
    fake = FakeVecEnv(2, horizon=3)
    env = RecordEpisodeVals(GymVectorEnv(venv=fake), buffer_length=10)
    env.reset()
    powers = []
    for k in range(6):
        _, r, _, trunc, infos = env.step(np.zeros((2, 3), dtype=np.float32))
        powers.append(infos["Power agent"].copy())
        if k in (2, 5):
            assert trunc.all() and infos["_episode"].all() and infos["episode"]["l"].tolist() == [3, 3]
    assert len(env.mean_power_queue) == 4 and list(env.length_queue) == [3, 3, 3, 3]
    assert env.mean_power_queue[0] == pytest.approx(np.mean([p[0] for p in powers[:3]]))
    assert env.return_queue[1] == pytest.approx(sum(p[1] for p in powers[:3]) / 1000.0, rel=1e-6)
    assert env.num_envs == 2 and env.episode_count == 4


def _loop_reference(ws, wd, ti, yaws, t_sim, baseline):
    """What eval_single_fast records for ONE condition on the fake env (serial reference of the batched assembly)."""
    env = FakeVecEnv(1, n_turb=3, baseline=baseline, eval_mode=True)
    env.set_wind_vals(ws=ws, ti=ti, wd=wd)
    agent = ConstantAgent(yaws)
    obs, _ = env.reset()
    pw, yw = [env.state["power"][0, 0].numpy().copy()], [env.state["yaw"][0, 0].numpy().copy()]
    rew = [0.0]
    for _ in range(1, t_sim):
        obs, r, _, _, _ = env.step(batch_actions(agent, obs, env))
        pw.append(env.state["power"][0, 0].numpy().copy()); yw.append(env.state["yaw"][0, 0].numpy().copy()); rew.append(float(r[0]))
    return np.array(pw), np.array(yw), np.array(rew)


@pytest.mark.parametrize("baseline", [False, True])
def test_eval_batched_matches_serial_loop_and_reference_layout(baseline):
    wss, wds, tis, t_sim = [8.0, 10.0, 12.0], [265, 270], [0.05], 6
    conds = condition_grid(wss, wds, tis, ["Default"])
    assert conds[1] == (8.0, 270, 0.05, "Default") and len(conds) == 6
    env = FakeVecEnv(len(conds), n_turb=3, baseline=baseline, eval_mode=True)
    ds = eval_batched(env, ConstantAgent([-10, 20, 0]), wss, wds, tis, t_sim=t_sim, model_step=7)
    assert ds["powerF_a"].shape == (t_sim, 3, 2, 1, 1, 1) and ds.dims("powerT_a")[1] == "turb"
    assert ds["powerT_a"].shape == (t_sim, 3, 3, 2, 1, 1, 1) and ds.coords["model_step"].tolist() == [7]
    assert ("pct_inc" in ds) == baseline and ("powerF_b" in ds) == baseline
    for i, ws in enumerate(wss):
        for j, wd in enumerate(wds):
            pw, yw, rew = _loop_reference(ws, wd, 0.05, [-10, 20, 0], t_sim, baseline)
            assert np.allclose(ds["powerT_a"][:, :, i, j, 0, 0, 0], pw)
            assert np.allclose(ds["yaw_a"][:, :, i, j, 0, 0, 0], yw)
            assert np.allclose(ds["powerF_a"][:, i, j, 0, 0, 0], pw.sum(1))
            assert np.allclose(ds["reward"][:, i, j, 0, 0, 0], rew)
            assert ds.coords["time0"][i, j, 0, 0] == 100 + int(ws)
    assert ds["yaw_a"][-1, :, 0, 0, 0, 0, 0].tolist() == [-5.0, 5.0, 0.0]   # one degree per step towards the target
    if baseline:
        assert np.allclose(ds["yaw_b"], 0.0) and ds["pct_inc"][0].max() == pytest.approx(0.0, abs=1e-4)


def test_eval_dataset_roundtrip_and_agent_eval(tmp_path):
    made = []

    def factory(n):
        made.append(FakeVecEnv(n, n_turb=3, eval_mode=True))
        return made[-1]
    ae = AgentEval(env_factory=factory, model=ConstantAgent([0, 0, 0]), name=str(tmp_path / "run"), t_sim=4)
    ae.set_conditions(winddirs=[260, 270, 280], windspeeds=[9, 11], turbintensities=[0.05, 0.1])
    ds = ae.eval_multiple()
    assert made[-1].n_envs == 12 and made[-1].closed
    assert ds["powerF_a"].shape == (4, 2, 3, 2, 1, 1)
    ae.save_performance()
    back = EvalDataset.load(str(tmp_path / "run_eval.npz"))
    assert np.array_equal(back["powerT_a"], ds["powerT_a"]) and back.dims("powerT_a")[1] == "turb"
    assert np.array_equal(back.coords["wd"], [260, 270, 280])
    ae.set_condition(ws=7.0, wd=255)
    one = ae.eval_single()
    assert one["powerF_a"].shape == (4, 1, 1, 1, 1, 1) and made[-1].n_envs == 1
    with pytest.raises(ValueError):
        eval_batched(FakeVecEnv(3), ConstantAgent([0, 0, 0]), [8.0, 9.0], t_sim=2)
    with pytest.raises(NotImplementedError):
        eval_batched(FakeVecEnv(1), ConstantAgent([0, 0, 0]), turbboxes=["box7"], t_sim=2)


def _ppo_golden():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_golden.npz"))
    sd = {k[3:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("sd/")}
    return z, sd


def test_shipped_ppo_agent_actions_match_the_numpy_evaluation():
    """The reference ships examples/PPO_2975000.zip (SB3 MlpPolicy 8 -> 64 -> 64 -> 4).  Its actor weights and its
    deterministic actions on 67 fixed observations (evaluated with plain numpy by tests/golden/make_golden.py:make_ppo)
    are committed; SB3MlpPolicy must reproduce them."""
    z, sd = _ppo_golden()
    pol = SB3MlpPolicy.from_state_dict(sd)
    assert [m.in_features for m in pol.policy_net if hasattr(m, "in_features")] == [8, 64] and pol.action_net.out_features == 4
    got = pol.predict_batch(torch.as_tensor(z["obs"])).numpy()
    assert got.shape == (67, 4) and np.abs(got - z["actions"]).max() < 2e-6
    assert np.abs(z["actions"]).max() <= 1.0 and np.std(z["actions"]) > 0.05      # a trained policy, not a constant


@pytest.mark.reference
def test_shipped_ppo_zip_loads_with_from_zip():
    """SB3MlpPolicy.from_zip on the reference's own model file (build container only)."""
    import os
    path = "/root/reference/examples/PPO_2975000.zip"
    if not os.path.isfile(path):
        pytest.skip("reference checkout not present")
    z, sd = _ppo_golden()
    pol = SB3MlpPolicy.from_zip(path)
    for k, v in sd.items():
        own = "policy_net." + k[len("mlp_extractor.policy_net."):] if k.startswith("mlp_extractor.") else k
        assert torch.equal(pol.state_dict()[own], v), k
    assert np.abs(pol.predict_batch(torch.as_tensor(z["obs"])).numpy() - z["actions"]).max() < 2e-6
    with pytest.raises(ValueError):
        import tempfile, zipfile
        with tempfile.NamedTemporaryFile(suffix=".zip") as f:
            with zipfile.ZipFile(f.name, "w") as zz:
                zz.writestr("data", "{}")
            SB3MlpPolicy.from_zip(f.name)


def test_lazy_infos_compute_power_sums_on_first_read():
    """as_torch adapters hand out the env's live info views; the farm power sums cost a reduction each and are only
    evaluated when somebody reads them (RecordEpisodeVals does, a bare rollout loop does not)."""
    from windgym_b200.vector import _LazyInfos
    p = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    raw = {"Power pr turbine agent": p, "Power pr turbine baseline": 2 * p, "Wind speed Global": np.ones(3)}
    inf = _LazyInfos(raw, baseline=True)
    assert "Power agent" in inf and "Power baseline" in inf and "nope" not in inf
    assert not dict.__contains__(inf, "Power agent")                   # not computed yet
    assert torch.equal(inf["Power agent"], p.sum(dim=1)) and dict.__contains__(inf, "Power agent")
    assert set(inf.keys()) == set(raw) | {"Power agent", "Power baseline"}
    assert torch.equal(dict(inf.items())["Power baseline"], 2 * p.sum(dim=1))
    with pytest.raises(KeyError):
        inf["nope"]
    assert "Power baseline" not in _LazyInfos(raw, baseline=False)
    fake = FakeVecEnv(4, n_turb=3, horizon=50)
    env = GymVectorEnv(venv=fake, as_torch=True)
    env.reset()
    _, _, _, trunc, infos = env.step(torch.zeros((4, 3)))
    assert trunc.dtype == torch.bool and infos["Power agent"].shape == (4,)
