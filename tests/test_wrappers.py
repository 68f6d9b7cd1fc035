"""Wrappers: numpy oracle pinned to the reference's wrapper classes (CPU), and the fused CUDA
kernels checked against the golden vectors and the oracle (GPU)."""
import glob
import json
import os

import numpy as np
import pytest

import util
from oracle import wrappers_oracle as WO
from gym_pcgrl_b200 import PROBLEMS, _abi

WRAPS = sorted(glob.glob(os.path.join(util.GOLDEN, "wrap_*.npz")))


def _load(path):
    d = np.load(path)
    return {k: d[k] for k in d.files if k != "meta"}, json.loads(str(d["meta"]))


def _replay_oracle_env(meta):
    """Underlying env stepped by the C oracle with the golden seeds (single env, manual reset)."""
    import oracle
    env = util.host_env(meta["env_id"], meta["kwargs"], num_envs=1, auto_reset=False)
    # the composite wrappers call adjust_param ONCE (wrappers.py:218,237); host_env calls it twice only if kwargs
    o = oracle.OracleEnv(env.native_config, 1)
    o.set_rng_states(util.randomstate_words(meta["seed"])[None])
    return env, o


@pytest.mark.parametrize("path", WRAPS, ids=[os.path.basename(p)[5:-4] for p in WRAPS])
def test_wrapper_oracle_matches_reference(path):
    g, meta = _load(path)
    assert not meta["kwargs"]
    env, o = _replay_oracle_env(meta)
    prob, rep, _ = meta["env_id"].split("-")
    T, binary = env.get_num_tiles(), prob == "binary"
    h, w = env._prob._height, env._prob._width
    border = env.get_border_tile()

    def image():
        m = o["map"][0]
        if meta["kind"] == "cropped":
            return WO.cropped_image(m, o["pos"][0], meta["crop"], border, T, binary)
        return WO.full_image(m, T, binary)

    o.reset()
    np.testing.assert_array_equal(image().astype(np.uint8), g["reset_obs"][0])
    k = 0
    for t in range(meta["steps"]):
        a = g["actions"][t]
        if meta["kind"] == "cropped":
            act = a
        else:
            act = WO.action_map(a[0], o["map"][0], None if rep == "wide" else o["pos"][0], h, w, T)
        o.step(np.asarray(act, np.int32).reshape(1, -1))
        np.testing.assert_array_equal(image().astype(np.uint8), g["obs"][t], err_msg="%s step %d" % (meta["name"], t))
        assert o["reward"][0] == g["reward"][t] and bool(o["done"][0]) == bool(g["done"][t])
        if g["done"][t]:
            o.reset()
            k += 1
            np.testing.assert_array_equal(image().astype(np.uint8), g["reset_obs"][k])


@pytest.mark.gpu
@pytest.mark.parametrize("path", WRAPS, ids=[os.path.basename(p)[5:-4] for p in WRAPS])
@pytest.mark.parametrize("out_dtype", ["uint8", "float32"])
def test_wrapper_cuda_matches_reference_golden(path, out_dtype):
    import torch
    from gym_pcgrl_b200 import wrappers as W
    g, meta = _load(path)
    if meta["kind"] == "cropped":
        env = W.CroppedImagePCGRLWrapper(meta["env_id"], meta["crop"], num_envs=1, out_dtype=out_dtype,
                                         env_kwargs=dict(auto_reset=False))
    else:
        env = W.ActionMapImagePCGRLWrapper(meta["env_id"], num_envs=1, out_dtype=out_dtype, env_kwargs=dict(auto_reset=False))
    env.pcgrl_env.set_rng_states(util.randomstate_words(meta["seed"])[None])
    assert list(env.shape) == meta["obs_shape"]
    obs = env.reset()
    np.testing.assert_array_equal(obs[0].cpu().numpy().astype(np.uint8), g["reset_obs"][0])
    k = 0
    for t in range(meta["steps"]):
        a = g["actions"][t]
        obs, r, d, info = env.step(torch.from_numpy(a[None] if len(a) > 1 else a))
        np.testing.assert_array_equal(obs[0].cpu().numpy().astype(np.uint8), g["obs"][t], err_msg="%s step %d" % (meta["name"], t))
        assert float(r[0]) == g["reward"][t] and bool(d[0]) == bool(g["done"][t])
        if g["done"][t]:
            obs = env.reset()
            k += 1
            np.testing.assert_array_equal(obs[0].cpu().numpy().astype(np.uint8), g["reset_obs"][k])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [("zelda-narrow-v0", 22), ("binary-turtle-v0", 28), ("sokoban-narrow-v0", 3), ("mdungeon-turtle-v0", 64)])
def test_batched_cropped_image_matches_oracle(case):
    """512 envs, auto-reset rollouts: every crop of the batch equals the numpy oracle's crop of that env's map."""
    import torch
    from gym_pcgrl_b200 import wrappers as W
    env_id, crop = case
    n = 512
    env = W.CroppedImagePCGRLWrapper(env_id, crop, num_envs=n)
    env.pcgrl_env.set_rng_states(np.stack([util.randomstate_words(i) for i in range(n)]))
    obs = env.reset()
    prob = env_id.split("-")[0]
    T, binary, border = env.get_num_tiles(), prob == "binary", env.get_border_tile()
    arng = np.random.RandomState(1)
    for t in range(12):
        a = arng.randint(env.action_space.n, size=n).astype(np.int32)
        obs, r, d, info = env.step(torch.from_numpy(a).cuda())
        maps, pos, got = env.pcgrl_env._tens["map"].cpu().numpy(), env.pcgrl_env._tens["pos"].cpu().numpy(), obs.cpu().numpy()
        for i in range(0, n, 7):
            np.testing.assert_array_equal(got[i], WO.cropped_image(maps[i], pos[i], crop, border, T, binary).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("prob,w,h,crop", [("binary", 16, 16, 28), ("binary", 13, 7, 28), ("binary", 9, 31, 12), ("binary", 32, 32, 64),
                                           ("binary", 5, 3, 4), ("binary", 16, 16, 0), ("binary", 12, 5, 0), ("zelda", 11, 7, 22),
                                           ("zelda", 11, 16, 22), ("mdungeon", 7, 11, 64), ("zelda", 9, 9, 0), ("mdungeon", 3, 5, 7)])
def test_uint8_fast_path_equals_float_path(prob, w, h, crop):
    """pcgrl_obs_image, uint8 output (warp-per-env word builder: funnel-shifted row windows for the raw index, one word per
    four one-hot channels) against the float32 output of the generic kernel on the same random maps and cursors, at ragged
    sizes (env strides that are not multiples of 4, crops wider than the map, cursors in every corner).  Exact equality."""
    import ctypes as C
    import torch
    from gym_pcgrl_b200 import PROBLEMS, REPRESENTATIONS, _native
    from gym_pcgrl_b200._config import build_config
    p = PROBLEMS[prob]()
    p.adjust_param(width=w, height=h)
    cfg = build_config(p, REPRESENTATIONS["narrow"](), 1, 1, auto_reset=False)
    T, n = len(p.tile_types), 301
    g = torch.Generator().manual_seed(w * 100 + h)
    maps = torch.randint(0, T, (n, h, w), generator=g, dtype=torch.uint8).cuda()
    pos = torch.stack([torch.randint(0, w, (n,), generator=g), torch.randint(0, h, (n,), generator=g)], dim=1).to(torch.uint8)
    pos[0] = torch.tensor([0, 0]); pos[1] = torch.tensor([w - 1, h - 1]); pos[2] = torch.tensor([0, h - 1]); pos[3] = torch.tensor([w - 1, 0])
    pos = pos.cuda()
    one_hot = prob != "binary"
    sh, sw, c = (crop or h), (crop or w), (T if one_hot else 1)
    outs = []
    for dt, code in ((torch.uint8, 0), (torch.float32, 1)):
        out = torch.full((n, sh, sw, c), 99, dtype=dt, device="cuda")
        rc = _native.lib().pcgrl_obs_image(C.byref(cfg), maps.data_ptr(), pos.data_ptr(), out.data_ptr(), n, crop, 1, int(one_hot), code,
                                           _native.stream_ptr(torch.device("cuda", 0)))
        assert rc == 0, _native.lib().pcgrl_last_error()
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0].float(), outs[1])
    # and one env against numpy directly
    m = maps[1].cpu().numpy()
    padded = np.pad(m, crop // 2, constant_values=1) if crop else m
    x, y = int(pos[1, 0]), int(pos[1, 1])
    win = padded[y:y + crop, x:x + crop] if crop else padded
    want = np.eye(T, dtype=np.uint8)[win] if one_hot else win[..., None]
    np.testing.assert_array_equal(outs[0][1].cpu().numpy(), want)


@pytest.mark.gpu
def test_batched_action_map_matches_oracle():
    import torch
    from gym_pcgrl_b200 import wrappers as W
    for env_id in ("zelda-wide-v0", "sokoban-narrow-v0", "binary-turtle-v0"):
        n = 256
        env = W.ActionMapImagePCGRLWrapper(env_id, num_envs=n)
        env.pcgrl_env.set_rng_states(np.stack([util.randomstate_words(i) for i in range(n)]))
        env.reset()
        rep = env_id.split("-")[1]
        arng = np.random.RandomState(2)
        for t in range(10):
            flat = arng.randint(env.action_space.n, size=n).astype(np.int32)
            maps, pos = env.pcgrl_env._tens["map"].cpu().numpy(), env.pcgrl_env._tens["pos"].cpu().numpy()
            want = [WO.action_map(flat[i], maps[i], None if rep == "wide" else pos[i], env.h, env.w, env.dim) for i in range(n)]
            obs, r, d, info = env.step(torch.from_numpy(flat))
            got = env._actions.cpu().numpy().reshape(n, -1)
            np.testing.assert_array_equal(got, np.asarray(want, np.int32).reshape(n, -1))
            T, binary = env.get_num_tiles(), env_id.startswith("binary")
            m2 = env.pcgrl_env._tens["map"].cpu().numpy()
            np.testing.assert_array_equal(obs[5].cpu().numpy(), WO.full_image(m2[5], T, binary).astype(np.uint8))
