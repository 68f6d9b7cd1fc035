"""Row f4 on the GPU: batched rendering against a numpy restatement of Problem.render + the cursor frame, and the
device-resident PPO consumer (a few updates end to end: fused env step -> fused wrappers -> policy -> PPO update)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _numpy_render(m, pos, atlas, bw, bh, border_tile, ts):
    """probs/problem.py:134-156 + reps/narrow_rep.py:126-140 + convert('RGB') for one map."""
    h, w = m.shape
    img = np.zeros(((h + 2 * bh) * ts, (w + 2 * bw) * ts, 4), np.uint8)
    for y in range(h + 2 * bh):
        for x in range(w + 2 * bw):
            inside = bw <= x < bw + w and bh <= y < bh + h
            tile = m[y - bh, x - bw] if inside else border_tile
            img[y * ts:(y + 1) * ts, x * ts:(x + 1) * ts] = atlas[tile]
    if pos is not None:
        x0, y0 = (int(pos[0]) + bw) * ts, (int(pos[1]) + bh) * ts
        frame = np.zeros((ts, ts), bool)
        frame[:2, :] = frame[-2:, :] = frame[:, :2] = frame[:, -2:] = True
        img[y0:y0 + ts, x0:x0 + ts][frame] = (255, 0, 0, 255)
    return img[:, :, :3]


@pytest.mark.parametrize("env_id,kwargs", [("zelda-narrow-v0", {}), ("binary-wide-v0", dict(width=9, height=5)), ("smb-turtle-v0", {})])
def test_batched_render_matches_numpy(env_id, kwargs):
    import torch
    n = 6
    env = util.host_env(env_id, kwargs, num_envs=n, device="cuda")
    env.reset()
    prob = env._prob
    rs = np.random.RandomState(0)
    atlas = rs.randint(0, 256, size=(len(prob.get_tile_types()), prob._tile_size, prob._tile_size, 4)).astype(np.uint8)
    prob._graphics = atlas
    img = env.render("rgb_array")
    maps, pos = env._tens["map"].cpu().numpy(), env._tens["pos"].cpu().numpy()
    bw, bh = prob._border_size
    assert tuple(img.shape) == (n, (prob._height + 2 * bh) * 16, (prob._width + 2 * bw) * 16, 3) and img.dtype == torch.uint8
    for i in range(n):
        want = _numpy_render(maps[i], None if env._rep.name == "wide" else pos[i], atlas, bw, bh, env.get_border_tile(), 16)
        np.testing.assert_array_equal(img[i].cpu().numpy(), want, err_msg="%s env %d" % (env_id, i))
    # default graphics: the grey levels of the reference's base class
    prob._graphics = None
    g = prob.get_graphics()
    assert g.shape == atlas.shape and g[1, 0, 0, 0] == int(255 / len(prob.get_tile_types())) and (g[:, :, :, 3] == 255).all()


@pytest.mark.parametrize("game,rep,envs", [("binary", "narrow", 256), ("zelda", "wide", 64), ("sokoban", "turtle", 64)])
def test_ppo_consumer_runs_on_device(game, rep, envs):
    import torch
    from gym_pcgrl_b200.ppo import PPO, make_training_env
    env = make_training_env(game, rep, envs)
    ppo = PPO(env, n_steps=16, nminibatches=2, noptepochs=2)
    lines = []
    ppo.learn(3 * 16 * envs, log=lines.append)
    assert len(lines) == 3 and ppo.num_timesteps == 3 * 16 * envs
    assert all(torch.isfinite(p).all() for p in ppo.policy.parameters())
    assert ppo.obs_buf.dtype == torch.uint8 and ppo.obs_buf.is_cuda
    env.pcgrl_env.check_status()


def test_ppo_save_load_predict_and_inference_tool(tmp_path):
    """model.save / PPO2.load / agent.predict of the reference's train.py + inference.py: a saved learner reloads to the same
    logits (torch modules and native kernels), and tools/inference.py plays one episode per env with it."""
    import importlib.util
    import os
    import torch
    from gym_pcgrl_b200.ppo import PPO, make_training_env
    env = make_training_env("binary", "narrow", 64)
    agent = PPO(env, n_steps=8, native_policy=True)
    agent.learn(8 * 64)
    path = str(tmp_path / "agent.pt")
    agent.save(path)
    env2 = make_training_env("binary", "narrow", 64)
    other = PPO(env2, n_steps=8, native_policy=True, seed=123).load(path)
    obs = env2.reset()
    with torch.no_grad():
        a, b = agent.policy(obs)[0], other.policy(obs)[0]
    assert torch.equal(a, b)
    acts = other.predict(obs, deterministic=True)
    assert acts.dtype == torch.int32 and acts.shape == (64,) and int(acts.max()) < int(env2.action_space.n)
    spec = importlib.util.spec_from_file_location("inference_tool", os.path.join(os.path.dirname(__file__), "..", "tools", "inference.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    out = str(tmp_path / "maps.npz")
    ret = tool.infer("binary", "narrow", path, num_envs=32, change_percentage=0.4, out=out)
    assert ret.shape == (32,)
    import numpy as np
    z = np.load(out)
    assert z["maps"].shape == (32, 14, 14) and z["render"].shape[0] == 16 and z["render"].shape[-1] == 3
