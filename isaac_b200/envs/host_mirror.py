"""Host-side images of an env's stacked observations, for consumers that live on the CPU.

The reference's runner moves `obs` and `critic_obs` to the learner's device after every step
(algo/ppo/on_policy_runner.py:136); with `rl_device = cpu` that is the whole `[N, 615]` + `[N, 1050]` stack over PCIe,
27 MB per step at 4096 envs - the bound of the end-to-end path.  But a step's stack is the previous one shifted by one
frame (envs/custom/hector_env.py:246-254): only the newest frame (41 + 70 floats per env) is new.  The mirror keeps, in
pinned host memory, a ring of frame slots per env that the GPU appends to (`hb_env_mirror_frames`: 2-D DMA copies - or
kernel stores - of the newest frames straight into the pinned host rings; the first S - 1 slots of a ring of C are written
twice so that the last S frames are always contiguous), and hands out the stacked observations as strided VIEWS of the
rings: `[N, S*F]` tensors with a row pitch of `(C + S - 1) * F` floats, bit-equal to the device tensors, with nothing
copied or shifted on the host.  About 2.6 MB per step instead of 27 MB (4096 hector envs, C = S + 17).

    mirror = HostObservationMirror(env)          # images the current observations once (a full copy)
    obs, priv, rew, reset, extras = env.step(actions)
    ticket = mirror.update(obs, priv)            # async: enqueued behind the step on the current stream
    ...                                          # (more steps may be enqueued: up to `spare - 1` updates can be outstanding)
    host_obs, host_priv = mirror.views(ticket)   # waits for that update only; CPU tensors

The history of an env that a step reset is zero (reset_idx, hector_env.py:256-261).  The GPU never rewrites slots that an
earlier step's views still show; the mirror zeroes those rows on the host when `views()` is called for the step that
reset them - so the views of a step stay intact until `views()` of the next step (and for at most `spare - 1` updates).
`views()` must be called for tickets in order (skipped tickets are processed on the way).
"""
from __future__ import annotations

import torch

from .. import _lib


def window_start(a: int, c: int, s: int) -> int:
    """First slot of the contiguous run that holds the last s frames when the newest one sits in slot a of a ring of c
    (+ s - 1 duplicated head slots)."""
    return a - (s - 1) if a >= s - 1 else a + c - (s - 1)


def history_slots(k: int, c: int, s: int) -> list:
    """Every ring slot (both copies) that holds one of the s - 1 frames before frame number k - 1: what a reset zeroes."""
    slots = []
    for j in range(1, s):
        sl = (k - 1 - j) % c
        slots.append(sl)
        if sl < s - 1:
            slots.append(sl + c)
    return slots


class HostObservationMirror:
    def __init__(self, env, spare: int = 17, use_dma: bool = True):
        if spare < 2:
            raise ValueError("the rings need at least two slots more than the stack has frames")
        self._lib = _lib.load(check_device=True)
        self.env = env
        # frames by 2-D DMA copies (the copy engine) or by kernel stores into the mapped rings (the same speed end to end; one
        # history by each engine at the same time was measured too: 17.4 against 18.8 M env-steps/s - the small PCIe writes
        # are the shared bound)
        self.use_dma = use_dma
        cfg = env.cfg.env
        self.device = env.device
        n = self.num_envs = env.num_envs
        self._fa, self._sa = cfg.num_single_obs, cfg.frame_stack
        self._fb, self._sb = cfg.single_num_privileged_obs, cfg.c_frame_stack
        self._ca, self._cb = self._sa + spare, self._sb + spare
        # [N, C + S - 1, F] pinned rings (device-accessible under unified addressing); a larger C means fewer frames
        # written twice ((C + S - 1) / C of a frame per step on average)
        self._ring_a = torch.zeros(n, self._ca + self._sa - 1, self._fa).pin_memory()
        self._ring_b = torch.zeros(n, self._cb + self._sb - 1, self._fb).pin_memory()
        self._ring_ptr_a, self._ring_ptr_b = self._ring_a.data_ptr(), self._ring_b.data_ptr()
        self._depth = spare - 1          # updates that may be outstanding: their frames land in slots no live view shows
        self._reset_host = [torch.zeros(n, dtype=torch.bool).pin_memory() for _ in range(self._depth)]
        self._events = [torch.cuda.Event() for _ in range(self._depth)]
        self._k0 = 0                     # frames in the rings when ticket 0 was issued
        self._issued = 0                 # tickets handed out
        self._done = 0                   # tickets whose frames have landed and whose resets have been applied
        self._views = {}
        # average bytes a step sends (the first S - 1 of every C slots are written twice) + the reset flags
        self.bytes_per_update = n * 4 * (self._fa * (self._ca + self._sa - 1) / self._ca + self._fb * (self._cb + self._sb - 1) / self._cb) + n
        self.resync(env.get_observations(), env.get_privileged_observations())

    # ------------------------------------------------------------------ views
    def _window(self, ring, c, s, f, k):
        """The strided [N, s*f] view of the last s frames after k appended frames (k >= 1)."""
        start = window_start((k - 1) % c, c, s)
        return torch.as_strided(ring, (self.num_envs, s * f), ((c + s - 1) * f, 1), storage_offset=start * f)

    def _views_at(self, k):
        key = ((k - 1) % self._ca, (k - 1) % self._cb)
        v = self._views.get(key)
        if v is None:          # one pair of view objects per ring position, built once
            v = self._views[key] = (self._window(self._ring_a, self._ca, self._sa, self._fa, k),
                                    self._window(self._ring_b, self._cb, self._sb, self._fb, k))
        return v

    def _zero_history(self, ids, k):
        """Rows `ids` were reset by the step whose frame is number k - 1: the S - 1 frames before it become zeros (both copies)."""
        for ring, c, s in ((self._ring_a, self._ca, self._sa), (self._ring_b, self._cb, self._sb)):
            ring[ids[:, None], torch.tensor(history_slots(k, c, s))[None, :]] = 0.0

    def views(self, ticket: int | None = None):
        """(obs, privileged_obs) host tensors of the step behind `ticket` (default: the latest update): waits for its frames,
        applies the resets of every step up to it, returns the strided views."""
        if ticket is None:
            ticket = self._issued - 1
        if ticket < 0:
            return self._views_at(self._k0)
        if ticket >= self._issued:
            raise ValueError("no such ticket")
        if ticket < self._done - 1:
            raise ValueError("the views of that step have been superseded (views() goes forward in time)")
        while self._done <= ticket:
            t = self._done
            self._events[t % self._depth].synchronize()
            mask = self._reset_host[t % self._depth]
            if bool(mask.any()):
                self._zero_history(mask.nonzero().flatten(), self._k0 + t + 1)
            self._done = t + 1
        return self._views_at(self._k0 + ticket + 1)

    # ------------------------------------------------------------------ full image (construction, after reset())
    def resync(self, obs: torch.Tensor, priv: torch.Tensor):
        """Image the whole stacks once (a full device -> host copy): at construction and whenever the env's observations
        were rewritten outside step() (reset(), reset_idx()).  Outstanding tickets are dropped."""
        torch.cuda.synchronize(self.device)
        k = max(self._sa, self._sb)              # both histories count appended frames together
        for ring, c, s, f, t in ((self._ring_a, self._ca, self._sa, self._fa, obs), (self._ring_b, self._cb, self._sb, self._fb, priv)):
            frames = t.detach().to("cpu").reshape(self.num_envs, s, f)
            ring.zero_()
            for j in range(s):           # frame j of the stack = appended frame number k - s + j
                slot = (k - s + j) % c
                ring[:, slot] = frames[:, j]
                if slot < s - 1:
                    ring[:, slot + c] = frames[:, j]
        self._k0, self._issued, self._done = k, 0, 0
        return self._views_at(k)

    # ------------------------------------------------------------------ per step
    def update(self, obs: torch.Tensor, priv: torch.Tensor, reset_buf: torch.Tensor | None = None) -> int:
        """Append the newest frames of a step's observations (the tensors step() returned) to the host rings, on the current
        stream; returns the ticket to hand to views().  `reset_buf` (device, bool / uint8 [N]) defaults to the env's - pass a
        snapshot if the env steps on before this stream gets to copy it."""
        if self._issued - self._done >= self._depth:
            raise RuntimeError(f"{self._depth} updates are outstanding: call views() before updating again (or build the mirror with a larger `spare`)")
        rb = self.env.reset_buf if reset_buf is None else reset_buf
        st = torch.cuda.current_stream(self.device)
        t = self._issued
        k = self._k0 + t
        rc = self._lib.hb_env_mirror_frames(
            obs.data_ptr(), obs.stride(0), self._sa * self._fa, self._fa, priv.data_ptr(), priv.stride(0), self._sb * self._fb, self._fb,
            None, self.num_envs, self._ring_ptr_a, self._ca, k % self._ca, self._ring_ptr_b, self._cb, k % self._cb, int(self.use_dma),
            st.cuda_stream)
        if rc:
            _lib.check(rc, "hb_env_mirror_frames")
        self._reset_host[t % self._depth].copy_(rb, non_blocking=True)
        self._events[t % self._depth].record(st)
        self._issued = t + 1
        return t

    def synchronize(self):
        """Block until every update has landed and its resets are applied."""
        if self._issued:
            self.views(self._issued - 1)
