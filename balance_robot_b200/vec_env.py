"""Batched balance-robot environments behind the Gymnasium / SB3 VecEnv surface.

The reference trains through `gym.make(id)` -> Monitor -> SB3's DummyVecEnv(1) (sb_rl.py:500-517); the
VecEnv methods SB3's collect_rollouts calls are reset(), step_async()/step_wait() (= step()), plus
num_envs / observation_space / action_space and the get_attr family.  `BalanceVecEnv`
keeps those names and the auto-reset contract (terminal_observation, TimeLimit.truncated, Monitor's
"episode" record) while running all N robots in one CUDA launch per step.

Two output modes:
  * output="torch" (default): observations / rewards / dones are CUDA tensors that stay on the device;
    `infos` is an `InfoBatch` (tensor-valued, converts lazily to SB3's list-of-dicts on indexing).
  * output="numpy": the exact SB3 contract (numpy arrays + list of dicts) through brb_env_step_host,
    i.e. host buffers with the copies inside the call.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi, mjcf, model as model_mod, registry


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (gymnasium is not installable here)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self, rng: Optional[np.random.Generator] = None):
        rng = rng or np.random.default_rng()
        return rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


# RobotBaseEnv.py:50-54 and :74-85
OBSERVATION_SPACE = Box([-2 * np.pi, -2 * np.pi, -1.0, -1.0, -1.0, -1.0], [2 * np.pi, 2 * np.pi, 1.0, 1.0, 1.0, 1.0])
ACTION_SPACE = Box([-1.0, -1.0], [1.0, 1.0])


_NO_INFO: dict = {}


class LazyInfoList(list):
    """SB3's `infos`: a real list of dicts, one per env.  Envs that did not finish share one empty dict; the dicts of the
    finished envs (terminal_observation / TimeLimit.truncated / Monitor's episode record) are built on first access, so a
    step costs one pointer fill of Python work instead of O(#done) dict constructions when the consumer only looks at a
    few of them.  `idx` is the ascending array of finished env indices (np.flatnonzero)."""

    def __init__(self, n, idx, tobs, trunc, epr, epl, t):
        super().__init__([_NO_INFO] * n)
        self._built = []
        self.refill(idx, tobs, trunc, epr, epl, t)

    def refill(self, idx, tobs, trunc, epr, epl, t):
        """Re-use this list for another step: only the entries that were materialised are reset (BalanceVecEnv keeps two
        of these and alternates, like its pinned output buffers, so creating `infos` is O(1) per step, not O(num_envs))."""
        for i in self._built:
            list.__setitem__(self, i, _NO_INFO)
        self._built = []
        self._idx = idx
        self._left = len(idx)          # finished envs whose dict has not been built yet
        self._src = (tobs, trunc, epr, epl, t)
        return self

    def _lookup(self, i):
        """position of env i among the finished envs, or -1"""
        k = int(np.searchsorted(self._idx, i))
        return k if k < len(self._idx) and self._idx[k] == i else -1

    def _build(self, i, k):
        tobs, trunc, epr, epl, t = self._src
        d = {"terminal_observation": np.array(tobs[k]), "TimeLimit.truncated": bool(trunc[k]),
             "episode": {"r": float(epr[k]), "l": int(epl[k]), "t": t}}
        list.__setitem__(self, i, d)
        self._built.append(i)
        self._left -= 1
        return d

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            i = int(i)
            if i < 0:
                i += len(self)
            d = list.__getitem__(self, i)
            if d is _NO_INFO and self._left:
                k = self._lookup(i)
                if k >= 0:
                    return self._build(i, k)
            return d
        if self._left:
            self._materialise()
        return list.__getitem__(self, i)

    def _materialise(self):
        if self._left:
            for k, i in enumerate(self._idx):
                if list.__getitem__(self, int(i)) is _NO_INFO:
                    self._build(int(i), k)

    def __iter__(self):
        self._materialise()
        return list.__iter__(self)

    def __eq__(self, other):
        self._materialise()
        return list.__eq__(self, other)


class InfoBatch:
    """Tensor-valued infos of one step; `infos[i]` materialises the SB3 dict for env i."""

    def __init__(self, done, truncated, terminal_obs, ep_return, ep_len, t_start):
        self.done, self.truncated, self.terminal_observation = done, truncated, terminal_obs
        self.episode_return, self.episode_length, self._t0 = ep_return, ep_len, t_start

    def __len__(self):
        return self.done.shape[0]

    def __getitem__(self, i: int) -> dict:
        if not bool(self.done[i]):
            return {}
        return {
            "terminal_observation": self.terminal_observation[i].detach().cpu().numpy(),
            "TimeLimit.truncated": bool(self.truncated[i]),
            "episode": {"r": float(self.episode_return[i]), "l": int(self.episode_length[i]),
                        "t": round(time.time() - self._t0, 6)},
        }


class BalanceVecEnv:
    metadata = {"render_modes": ["rgb_array"], "render_fps": 200}   # RobotBaseEnv.py:30-37 (no renderer here)

    def __init__(self, env_id: str, num_envs: int, device="cuda:0", seed: int = 0, env_id_offset: int = 0,
                 output: str = "torch", actderiv_skip_clamped: bool = True, truncate_unsupported: bool = False,
                 wheel_block: bool = False):
        if output not in ("torch", "numpy"):
            raise ValueError("output must be 'torch' or 'numpy'")
        self.spec = registry.spec(env_id)
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.BrbError("BalanceVecEnv runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise _cabi.BrbError("no CUDA device available (no CPU fallback)")
        self.output = output
        self.observation_space, self.action_space = OBSERVATION_SPACE, ACTION_SPACE
        self.render_mode = None
        self.robot = model_mod.compile_model(mjcf.parse(self.spec.scene), self.spec.kind, self.spec.max_episode_steps,
                                             actderiv_skip_clamped=actderiv_skip_clamped, truncate_unsupported=truncate_unsupported,
                                             wheel_block=wheel_block)
        L = _cabi.lib()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._model = C.c_void_p()
        tt = np.ascontiguousarray(self.robot.time_table, np.float64)
        _cabi.check(L.brb_model_create(C.byref(self.robot.consts), tt.ctypes.data, len(tt), dev_index, C.byref(self._model)),
                    "brb_model_create")
        self._env = C.c_void_p()
        self._seed, self._env_id_offset, self._reseed = int(seed), int(env_id_offset), None
        _cabi.check(L.brb_env_create(self._model, self.num_envs, seed, env_id_offset, C.byref(self._env)), "brb_env_create")
        n, dv = self.num_envs, self.device
        self._obs = torch.zeros((n, 6), dtype=torch.float32, device=dv)
        self._rew = torch.zeros(n, dtype=torch.float32, device=dv)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dv)
        self._trunc = torch.zeros(n, dtype=torch.uint8, device=dv)
        self._tobs = torch.zeros((n, 6), dtype=torch.float32, device=dv)
        self._epr = torch.zeros(n, dtype=torch.float32, device=dv)
        self._epl = torch.zeros(n, dtype=torch.int32, device=dv)
        if output == "numpy":
            pin = dict(pin_memory=True)
            self._h_act = torch.zeros((n, 2), dtype=torch.float32, **pin)
            # obs | reward | done of a buffer set live in ONE pinned block laid out like the device staging block, so that they
            # come back in a single device-to-host copy; the finished-episode rows are written into `rows` by the device directly
            offs, total = (C.c_int64 * 3)(), C.c_int64()
            _cabi.check(L.brb_env_host_layout(self._env, C.byref(offs), C.byref(total)), "brb_env_host_layout")
            self._hbuf = []
            for _ in range(2):
                blk = torch.zeros(total.value, dtype=torch.uint8, **pin)
                hb = dict(block=blk, obs=blk[offs[0]:offs[0] + 24 * n].view(torch.float32).view(n, 6),
                          rew=blk[offs[1]:offs[1] + 4 * n].view(torch.float32), done=blk[offs[2]:offs[2] + n],
                          rows=torch.zeros((n, _cabi.DONE_ROW_WORDS), dtype=torch.float32, **pin))
                # numpy views and raw addresses are made once: a step is ~0.9 ms, every microsecond of Python shows
                hb["np"] = (hb["obs"].numpy(), hb["rew"].numpy(), hb["done"].numpy().view(np.bool_), hb["rows"].numpy())
                hb["ptr"] = (hb["obs"].data_ptr(), hb["rew"].data_ptr(), hb["done"].data_ptr(), hb["rows"].data_ptr())
                self._hbuf.append(hb)
            self._h_ndone = torch.zeros(1, dtype=torch.int32, **pin)
            self._h_act_np, self._h_ndone_np = self._h_act.numpy(), self._h_ndone.numpy()
            self._h_ptr = (self._h_act.data_ptr(), self._h_ndone.data_ptr())
            self._flip = 0
        self._actions = None
        self._t0 = time.time()
        self.reset_infos: List[dict] = [{} for _ in range(min(n, 1))]
        self._closed = False

    # ------------------------------------------------------------------ VecEnv API
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def seed(self, seed: Optional[int] = None):
        """SB3 VecEnv.seed: re-keys the Philox streams; takes effect at the next reset() (the env object is rebuilt)."""
        self._reseed = seed
        return [seed] * self.num_envs

    def _apply_reseed(self) -> None:
        seed, self._reseed = self._reseed, None
        if seed is None:
            return
        L = _cabi.lib()
        L.brb_env_destroy(self._env)
        self._env = C.c_void_p()
        self._seed = int(seed)
        _cabi.check(L.brb_env_create(self._model, self.num_envs, self._seed, self._env_id_offset, C.byref(self._env)), "brb_env_create")

    def reset(self, replay_u: Optional[torch.Tensor] = None):
        L = _cabi.lib()
        self._apply_reseed()
        ru = None
        if replay_u is not None:
            ru = replay_u.to(self.device, torch.float64).contiguous()
            assert ru.shape == (self.num_envs, 32 if self.spec.kind == 3 else 16)
        with torch.cuda.device(self.device):
            _cabi.check(L.brb_env_reset_all(self._env, self._obs.data_ptr(), ru.data_ptr() if ru is not None else None,
                                            self._stream()), "brb_env_reset_all")
        if self.output == "numpy":
            return self._obs.cpu().numpy()
        return self._obs.clone()

    def step_async(self, actions) -> None:
        self._actions = actions

    def step_wait(self, replay_u: Optional[torch.Tensor] = None):
        L = _cabi.lib()
        if self.output == "numpy":
            np.copyto(self._h_act_np, np.asarray(self._actions).reshape(self.num_envs, 2), casting="unsafe")
            self._flip ^= 1
            hb = self._hbuf[self._flip]
            p_obs, p_rew, p_done, p_rows = hb["ptr"]
            # finished-episode records arrive compacted (ascending env index): only obs / reward / done are full arrays
            rc = L.brb_env_step_host_compact(self._env, self._h_ptr[0], p_obs, p_rew, p_done, self._h_ptr[1], p_rows, self.num_envs)
            if rc:
                _cabi.check(rc, "brb_env_step_host_compact")
            np_obs, np_rew, np_done, np_rows = hb["np"]
            rows = np_rows[:int(self._h_ndone_np[0])]
            ints = rows.view(np.int32)
            src = (ints[:, 0], rows[:, 1:7], ints[:, 9], rows[:, 7], ints[:, 8], round(time.time() - self._t0, 6))
            if "infos" in hb:
                infos = hb["infos"].refill(*src)
            else:
                infos = hb["infos"] = LazyInfoList(self.num_envs, *src)
            # the returned arrays (and the infos list) belong to this step's buffer set; the other set is used by the next
            # step, so they stay valid for one more step() (SB3 copies them into its rollout buffer right away)
            return np_obs, np_rew, np_done, infos
        a = self._actions
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a, dtype=np.float32))
        a = a.to(self.device, torch.float32).contiguous()
        assert a.shape == (self.num_envs, 2), a.shape
        ru = None
        if replay_u is not None:
            ru = replay_u.to(self.device, torch.float64).contiguous()
            assert ru.shape == (self.num_envs, 40 if self.spec.kind == 3 else 20)
        with torch.cuda.device(self.device):
            _cabi.check(L.brb_env_step(self._env, a.data_ptr(), self._obs.data_ptr(), self._rew.data_ptr(), self._done.data_ptr(),
                                       self._trunc.data_ptr(), self._tobs.data_ptr(), self._epr.data_ptr(), self._epl.data_ptr(),
                                       ru.data_ptr() if ru is not None else None, self._stream()), "brb_env_step")
        infos = InfoBatch(self._done, self._trunc, self._tobs, self._epr, self._epl, self._t0)
        return self._obs, self._rew, self._done, infos

    def step(self, actions, replay_u: Optional[torch.Tensor] = None):
        self.step_async(actions)
        return self.step_wait(replay_u)

    def close(self) -> None:
        if not self._closed:
            L = _cabi.lib()
            L.brb_env_destroy(self._env)
            L.brb_model_destroy(self._model)
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        return [getattr(self, attr_name)] * len(self._indices(indices))

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        return [getattr(self, method_name)(*args, **kwargs)] * len(self._indices(indices))

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(self._indices(indices))

    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else list(indices)

    # ------------------------------------------------------------------ state access (trajectory checks)
    def get_state(self):
        """(qpos [N,nq], qvel [N,nv], xquat [N,4]) fp64 CUDA tensors — MuJoCo's data.qpos / data.qvel / chassis xquat
        (nq, nv = 9, 8; Env03-v2 appends the block: 16, 14)."""
        n = self.num_envs
        qpos = torch.empty((n, self.robot.consts.nq), dtype=torch.float64, device=self.device)
        qvel = torch.empty((n, self.robot.consts.nv), dtype=torch.float64, device=self.device)
        xquat = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().brb_env_get_state(self._env, qpos.data_ptr(), qvel.data_ptr(), xquat.data_ptr(), self._stream()),
                        "brb_env_get_state")
        return qpos, qvel, xquat

    def set_state(self, qpos, qvel) -> None:
        """MujocoEnv.set_state (+ mj_forward): kinematics become fresh."""
        qpos = torch.as_tensor(qpos).to(self.device, torch.float64).contiguous()
        qvel = torch.as_tensor(qvel).to(self.device, torch.float64).contiguous()
        assert qpos.shape == (self.num_envs, self.robot.consts.nq) and qvel.shape == (self.num_envs, self.robot.consts.nv)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().brb_env_set_state(self._env, qpos.data_ptr(), qvel.data_ptr(), self._stream()), "brb_env_set_state")

    def elapsed_steps(self) -> torch.Tensor:
        out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().brb_env_get_elapsed(self._env, out.data_ptr(), self._stream()), "brb_env_get_elapsed")
        return out

    def stats(self) -> dict:
        out = (C.c_uint64 * _cabi.NSTATS)()
        _cabi.check(_cabi.lib().brb_env_get_stats(self._env, C.byref(out)), "brb_env_get_stats")
        return {k: int(v) for k, v in zip(_cabi.STAT_NAMES, out) if k != "_"}

    def num_launches(self) -> int:
        return int(_cabi.lib().brb_env_num_launches(self._env))


def make_vec(env_id: str, num_envs: int, device="cuda:0", seed: int = 0, env_id_offset: int = 0, output: str = "torch",
             **kw) -> BalanceVecEnv:
    """The batched analogue of `gym.make(env_id)` (reference sb_rl.py:500)."""
    return BalanceVecEnv(env_id, num_envs, device=device, seed=seed, env_id_offset=env_id_offset, output=output, **kw)
