"""Host-side images of an env's stacked observations, for consumers that live on the CPU.

The reference's runner moves `obs` and `critic_obs` to the learner's device after every step
(algo/ppo/on_policy_runner.py:136); with `rl_device = cpu` that is the whole `[N, 615]` + `[N, 1050]` stack over PCIe,
27 MB per step at 4096 envs - the bound of the end-to-end path.  But a step's stack is the previous one shifted by one
frame (envs/custom/hector_env.py:246-254): only the newest frame (41 + 70 floats per env) is new.  The mirror keeps, in
pinned host memory, a ring of frame slots per env that the GPU appends to (`hb_env_mirror_frames`: 2-D DMA copies - or
kernel stores - of the newest frames straight into the pinned host rings; the first S - 1 slots of a ring of C are written
twice so that the last S frames are always contiguous), and hands out the stacked observations as strided VIEWS of the rings: `[N, S*F]` tensors
with a row pitch of `(C + S - 1) * F` floats, bit-equal to the device tensors, with nothing copied or shifted on the host.
About 2.6 MB per step instead of 27 MB (4096 hector envs, C = S + 17).

    mirror = HostObservationMirror(env)          # images the current observations once (a full copy)
    obs, priv, rew, reset, extras = env.step(actions)
    host_obs, host_priv = mirror.update(obs, priv)      # async: enqueued behind the step on the current stream
    mirror.synchronize()                                 # then read host_obs / host_priv (valid until the next update)
"""
from __future__ import annotations

import torch

from .. import _lib


class HostObservationMirror:
    def __init__(self, env, spare: int = 17, use_dma: bool = True):
        if spare < 1:
            raise ValueError("the ring needs at least one slot more than the stack has frames")
        self._lib = _lib.load(check_device=True)
        self.env = env
        # frames by 2-D DMA copies (the copy engine: 232 us per e2e step at 4096 envs) or by kernel stores into the mapped
        # rings (261 us); either way the kernel zeroes the rings of the envs a step reset
        self.use_dma = use_dma
        cfg = env.cfg.env
        self.device = env.device
        n = self.num_envs = env.num_envs
        self._fa, self._sa = cfg.num_single_obs, cfg.frame_stack
        self._fb, self._sb = cfg.single_num_privileged_obs, cfg.c_frame_stack
        self._ca, self._cb = self._sa + spare, self._sb + spare
        # [N, C + S - 1, F] pinned rings (device-accessible under unified addressing): the kernel writes them over PCIe; a
        # larger C means fewer frames written twice ((C + S - 1) / C of a frame per step on average)
        self._ring_a = torch.zeros(n, self._ca + self._sa - 1, self._fa).pin_memory()
        self._ring_b = torch.zeros(n, self._cb + self._sb - 1, self._fb).pin_memory()
        self._ring_ptr_a, self._ring_ptr_b = self._ring_a.data_ptr(), self._ring_b.data_ptr()
        self._k = 0                      # frames appended so far
        self._views = {}
        self._event = torch.cuda.Event()
        # average bytes a step sends (the first S - 1 of every C slots are written twice)
        self.bytes_per_update = n * 4 * (self._fa * (self._ca + self._sa - 1) / self._ca + self._fb * (self._cb + self._sb - 1) / self._cb)
        self.resync(env.get_observations(), env.get_privileged_observations())

    # ------------------------------------------------------------------ views
    def _window(self, ring, c, s, f, k):
        """The strided [N, s*f] view of the last s frames after k appended frames (k >= 1)."""
        a = (k - 1) % c
        start = a - (s - 1) if a >= s - 1 else a + c - (s - 1)
        return torch.as_strided(ring, (self.num_envs, s * f), ((c + s - 1) * f, 1), storage_offset=start * f)

    def views(self):
        """(obs, privileged_obs) host views of the most recently mirrored step."""
        key = ((self._k - 1) % self._ca, (self._k - 1) % self._cb)
        v = self._views.get(key)
        if v is None:          # one pair of view objects per ring position, built once
            v = self._views[key] = (self._window(self._ring_a, self._ca, self._sa, self._fa, self._k),
                                    self._window(self._ring_b, self._cb, self._sb, self._fb, self._k))
        return v

    # ------------------------------------------------------------------ full image (construction, after reset())
    def resync(self, obs: torch.Tensor, priv: torch.Tensor):
        """Image the whole stacks once (a full device -> host copy): at construction and whenever the env's observations
        were rewritten outside step() (reset(), reset_idx())."""
        torch.cuda.current_stream(self.device).synchronize()
        self._k = max(self._sa, self._sb)        # both histories count appended frames together
        for ring, c, s, f, t in ((self._ring_a, self._ca, self._sa, self._fa, obs), (self._ring_b, self._cb, self._sb, self._fb, priv)):
            frames = t.detach().to("cpu").reshape(self.num_envs, s, f)
            ring.zero_()
            for j in range(s):           # frame j of the stack = appended frame number k - s + j
                slot = (self._k - s + j) % c
                ring[:, slot] = frames[:, j]
                if slot < s - 1:
                    ring[:, slot + c] = frames[:, j]
        return self.views()

    # ------------------------------------------------------------------ per step
    def update(self, obs: torch.Tensor, priv: torch.Tensor, reset_buf: torch.Tensor | None = None):
        """Append the newest frames of a step's observations (the tensors step() returned) to the host rings, on the current
        stream; returns the host views of this step.  `reset_buf` defaults to the env's (the envs the step just reset get
        their history zeroed, like the device stacks)."""
        rb = self.env.reset_buf if reset_buf is None else reset_buf
        st = torch.cuda.current_stream(self.device)
        k = self._k
        rc = self._lib.hb_env_mirror_frames(
            obs.data_ptr(), obs.stride(0), self._sa * self._fa, self._fa, priv.data_ptr(), priv.stride(0), self._sb * self._fb, self._fb,
            rb.data_ptr(), self.num_envs, self._ring_ptr_a, self._ca, k % self._ca, self._ring_ptr_b, self._cb, k % self._cb, int(self.use_dma),
            st.cuda_stream)
        if rc:
            _lib.check(rc, "hb_env_mirror_frames")
        self._k = k + 1
        self._event.record(st)
        return self.views()

    def synchronize(self):
        """Block until the last update() has landed in host memory."""
        self._event.synchronize()
