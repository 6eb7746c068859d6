"""TsidEngine — torch-tensor front of libtsidb.so (one handle = one model+conf on one GPU).

PyTorch is plumbing here: device memory, streams, and (in sharding.py) torch.distributed.
All arithmetic happens in the CUDA library; nothing in this module computes a tick on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import TsidbAuxOut, TsidbGaitConf, TsidbRefs, check, conf_to_c, load_library, model_to_c
from .model_compiler import CompiledModel

REF_KEYS = ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf", "posture")


class TickOutput:
    """Result of one batched tick: what `sol` + the decode calls give in the reference
    (ref:main.py:121-127), for N envs."""

    __slots__ = ("tau", "ddq", "f", "status", "iters", "active_set", "com", "foot_lf", "foot_rf", "wrench", "lam", "lam_row")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class TsidEngine:
    def __init__(self, model: CompiledModel, conf, lf_frame: str, rf_frame: str, legacy: bool = False,
                 max_envs: int = 65536, device: int = 0):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise _capi.TsidbError("no CUDA device: the TSID tick has no CPU fallback")
        self.model = model
        self.cm = model_to_c(model, lf_frame, rf_frame)
        self.cc = conf_to_c(conf, model, legacy=legacy)
        self.na, self.nv, self.nq = model.na, model.nv, model.nq
        self._ref_dims = tuple(zip(REF_KEYS, (9, 24, 24, 12, 12, model.na)))
        self.device = torch.device("cuda", device)
        self._dev_index = int(device)
        self._tick_cache: Dict[tuple, TickOutput] = {}
        self._refs_struct, self._aux_struct = TsidbRefs(), TsidbAuxOut()
        self._refs_ref, self._aux_ref = C.byref(self._refs_struct), C.byref(self._aux_struct)
        self.max_envs = int(max_envs)
        h = C.c_void_p()
        check(self.lib.tsidb_create(C.byref(self.cm), C.byref(self.cc), self.max_envs, device, C.byref(h)), "tsidb_create")
        self.h = h
        self._out_cache: Dict[int, dict] = {}

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.tsidb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        """Raw handle of torch's current stream on the engine's device (the private fast getter when torch has it: the
        public one builds a Stream object per call, 4 us of a 75 us single-robot tick)."""
        try:
            return torch._C._cuda_getCurrentRawStream(self._dev_index)
        except AttributeError:
            return torch.cuda.current_stream(self.device).cuda_stream

    def _chk(self, t: torch.Tensor, n: int, nd: int, name: str, dtype=torch.float64) -> torch.Tensor:
        try:
            ok = t.is_cuda and t.dtype is dtype and t.get_device() == self._dev_index
        except AttributeError:
            ok = False
        if not ok:
            raise TypeError(f"{name}: expected a {dtype} tensor on {self.device}")
        if t.shape != (n, nd) or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous [{n}, {nd}] tensor, got {tuple(t.shape)}")
        return t

    def _outputs(self, n: int, aux: bool) -> dict:
        o = self._out_cache.get(n)
        if o is None:
            f64 = dict(dtype=torch.float64, device=self.device)
            o = {
                "tau": torch.empty((n, self.na), **f64), "ddq": torch.empty((n, self.nv), **f64),
                "f": torch.empty((n, 24), **f64),
                "status": torch.empty(n, dtype=torch.int32, device=self.device),
                "iters": torch.empty(n, dtype=torch.int32, device=self.device),
                "active_set": torch.empty((3, n), dtype=torch.int64, device=self.device),
                "com": torch.empty((n, 9), **f64), "foot_lf": torch.empty((n, 12), **f64),
                "foot_rf": torch.empty((n, 12), **f64), "wrench": torch.empty((n, 12), **f64),
                "lam": torch.empty((n, 32), **f64), "lam_row": torch.empty((n, 32), dtype=torch.int32, device=self.device),
            }
            # a few sizes stay resident (a single-robot kinematics()/solve() call between two batched ticks must
            # not evict the batch's buffers); the least recently created goes first
            while len(self._out_cache) >= 4:
                old_n = next(iter(self._out_cache))
                self._out_cache.pop(old_n)
                for key in [k for k in self._tick_cache if k[0] == old_n]:
                    self._tick_cache.pop(key)
            self._out_cache[n] = o
        return o

    def _tick_output(self, o: dict, aux: bool, want_active: bool) -> TickOutput:
        """Fields the call did not write are None (never a stale or uninitialised buffer).  The view object is cached with
        the buffers it wraps (the same tensors come back for the same batch size anyway)."""
        key = (o["tau"].shape[0], aux, want_active)
        t = self._tick_cache.get(key)
        if t is None or t.tau is not o["tau"]:
            keys = ["tau", "ddq", "f", "status", "iters"] + (["active_set"] if want_active else []) + \
                   (["com", "foot_lf", "foot_rf", "wrench", "lam", "lam_row"] if aux else [])
            t = TickOutput(**{k: o[k] for k in keys})
            self._tick_cache[key] = t
        return t

    # ------------------------------------------------------------------ API
    def set_default_refs(self, refs: Dict[str, np.ndarray]) -> None:
        arrs = [np.ascontiguousarray(refs[k], dtype=np.float64) for k in REF_KEYS]
        sizes = (9, 24, 24, 12, 12, self.na)
        for a, s, k in zip(arrs, sizes, REF_KEYS):
            if a.shape != (s,):
                raise ValueError(f"default ref {k}: expected [{s}], got {a.shape}")
        ptrs = [a.ctypes.data_as(_capi.c_double_p) for a in arrs]
        check(self.lib.tsidb_set_default_refs(self.h, *ptrs), "tsidb_set_default_refs")

    def compute(self, q: torch.Tensor, v: torch.Tensor, contact_mask: Optional[torch.Tensor] = None,
                refs: Optional[Dict[str, torch.Tensor]] = None, aux: bool = False, want_active: bool = True) -> TickOutput:
        """One batched tick.  The returned tensors are the engine's cached output buffers for this batch size: the
        NEXT compute()/rollout()/kinematics() call with the same size overwrites them in place (clone what must
        survive).  aux=False leaves com/foot_lf/foot_rf/wrench None, want_active=False leaves active_set None.
        All calls on one engine must be ordered on one CUDA stream (the handle has one set of workspaces)."""
        n = q.shape[0]
        self._chk(q, n, self.nq, "q")
        self._chk(v, n, self.nv, "v")
        if contact_mask is not None:
            if contact_mask.dtype is not torch.uint8 or not contact_mask.is_cuda or contact_mask.get_device() != self._dev_index \
                    or contact_mask.shape != (n,):
                raise TypeError("contact_mask: expected a uint8 [N] tensor on the engine's device")
        # the argument structs are reused from call to call (a single-robot tick is ~75 us: every microsecond of
        # marshalling shows)
        r = self._refs_struct
        for k, nd in self._ref_dims:
            t = refs.get(k) if refs else None
            setattr(r, k, None if t is None else self._chk(t, n, nd, k).data_ptr())
        o = self._outputs(n, aux)
        if aux:
            a = self._aux_struct
            a.com, a.foot_lf, a.foot_rf, a.wrench = (o[k].data_ptr() for k in ("com", "foot_lf", "foot_rf", "wrench"))
            a.lambda_, a.lambda_row = o["lam"].data_ptr(), o["lam_row"].data_ptr()
        ptrs = o.get("_ptrs")
        if ptrs is None:
            ptrs = o["_ptrs"] = tuple(o[k].data_ptr() for k in ("tau", "ddq", "f", "status", "iters", "active_set"))
        check(self.lib.tsidb_compute(
            self.h, n, 0, q.data_ptr(), v.data_ptr(), contact_mask.data_ptr() if contact_mask is not None else None,
            self._refs_ref, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5] if want_active else None,
            self._aux_ref if aux else None, self._stream()), "tsidb_compute")
        return self._tick_output(o, aux, want_active)

    def host_buffers(self, n: int, pinned: bool = True) -> Dict[str, np.ndarray]:
        """Output buffers for :meth:`compute_host`.  Pinned buffers (the default) are DMA targets themselves;
        pageable ones go through the library's staging copy."""
        shapes = {"tau": ((n, self.na), torch.float64), "ddq": ((n, self.nv), torch.float64), "f": ((n, 24), torch.float64),
                  "status": ((n,), torch.int32), "iters": ((n,), torch.int32), "active_set": ((3, n), torch.int64)}
        out = {}
        self._pinned_keep = getattr(self, "_pinned_keep", [])
        for k, (shp, dt) in shapes.items():
            t = torch.empty(shp, dtype=dt, pin_memory=pinned)
            self._pinned_keep.append(t)
            a = t.numpy()
            out[k] = a.view(np.uint64) if k == "active_set" else a
        return out

    @staticmethod
    def pin(a: np.ndarray) -> np.ndarray:
        """A pinned copy of a host array (so that compute_host can DMA straight from it)."""
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        TsidEngine._pins.append(t)  # the numpy view does not own the pinned allocation
        return t.numpy()

    _pins: list = []

    def compute_host(self, q: np.ndarray, v: np.ndarray, contact_mask: Optional[np.ndarray] = None,
                     refs: Optional[Dict[str, np.ndarray]] = None, want_active: bool = True,
                     out: Optional[Dict[str, np.ndarray]] = None, tau_only: bool = False) -> Dict[str, np.ndarray]:
        """The same tick with HOST numpy buffers in and out (H2D, kernels, D2H inside the call, chunked so that
        the copies overlap the kernels).  `out` = buffers from :meth:`host_buffers` to avoid per-call allocation.
        tau_only=True brings back only tau, status and iters (what ref:main.py:126 sends to the actuators): ddq and f
        stay on the device and out["ddq"], out["f"] are left untouched."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = q.shape[0]
        if q.shape != (n, self.nq) or v.shape != (n, self.nv):
            raise ValueError("q/v: expected [N,nq] / [N,nv]")
        r = TsidbRefs()
        keep = []
        if refs:
            for k, nd in zip(REF_KEYS, (9, 24, 24, 12, 12, self.na)):
                if refs.get(k) is not None:
                    arr = np.ascontiguousarray(refs[k], dtype=np.float64)
                    if arr.shape != (n, nd):
                        raise ValueError(f"refs[{k}]: expected [{n},{nd}]")
                    keep.append(arr)
                    setattr(r, k, arr.ctypes.data)
        m = None
        if contact_mask is not None:
            m = np.ascontiguousarray(contact_mask, dtype=np.uint8)
        if out is None:
            out = self.host_buffers(n, pinned=False)
        elif out["tau"].shape[0] != n:
            raise ValueError("out: buffers were allocated for a different batch size")
        check(self.lib.tsidb_compute_host(
            self.h, n, q.ctypes.data, v.ctypes.data, m.ctypes.data if m is not None else None, C.byref(r),
            out["tau"].ctypes.data, None if tau_only else out["ddq"].ctypes.data, None if tau_only else out["f"].ctypes.data,
            out["status"].ctypes.data, out["iters"].ctypes.data, out["active_set"].ctypes.data if want_active else None),
            "tsidb_compute_host")
        return out

    def compute_host_devrefs(self, q: np.ndarray, v: np.ndarray, contact_mask: Optional[torch.Tensor] = None,
                             refs: Optional[Dict[str, torch.Tensor]] = None, want_active: bool = True,
                             out: Optional[Dict[str, np.ndarray]] = None, tau_only: bool = False) -> Dict[str, np.ndarray]:
        """compute_host with the references and the contact phases resident on the DEVICE (CUDA tensors, e.g. the
        views of gait_state()): only q and v cross PCIe on the way in (tsidb_compute_host_devrefs); tau_only as in
        :meth:`compute_host`."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = q.shape[0]
        if q.shape != (n, self.nq) or v.shape != (n, self.nv):
            raise ValueError("q/v: expected [N,nq] / [N,nv]")
        r = TsidbRefs()
        if refs:
            for k, nd in zip(REF_KEYS, (9, 24, 24, 12, 12, self.na)):
                t = refs.get(k)
                if t is not None:
                    setattr(r, k, self._chk(t, n, nd, f"refs[{k}]").data_ptr())
        if contact_mask is not None:
            if contact_mask.dtype != torch.uint8 or contact_mask.device != self.device or tuple(contact_mask.shape) != (n,):
                raise TypeError("contact_mask: expected a uint8 [N] tensor on the engine's device")
        if out is None:
            out = self.host_buffers(n, pinned=False)
        elif out["tau"].shape[0] != n:
            raise ValueError("out: buffers were allocated for a different batch size")
        torch.cuda.current_stream(self.device).synchronize()  # the device arrays must be final: the call uses its own streams
        check(self.lib.tsidb_compute_host_devrefs(
            self.h, n, q.ctypes.data, v.ctypes.data, contact_mask.data_ptr() if contact_mask is not None else None, C.byref(r),
            out["tau"].ctypes.data, None if tau_only else out["ddq"].ctypes.data, None if tau_only else out["f"].ctypes.data,
            out["status"].ctypes.data, out["iters"].ctypes.data, out["active_set"].ctypes.data if want_active else None),
            "tsidb_compute_host_devrefs")
        return out

    def kinematics(self, q: torch.Tensor, v: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(com9, foot_lf12, foot_rf12): robot.com / robot.framePosition without a solve."""
        n = q.shape[0]
        self._chk(q, n, self.nq, "q")
        if v is not None:
            self._chk(v, n, self.nv, "v")
        o = self._outputs(n, True)
        a = TsidbAuxOut()
        a.com, a.foot_lf, a.foot_rf = o["com"].data_ptr(), o["foot_lf"].data_ptr(), o["foot_rf"].data_ptr()
        check(self.lib.tsidb_kinematics(self.h, n, 0, q.data_ptr(), v.data_ptr() if v is not None else None,
                                        C.byref(a), self._stream()), "tsidb_kinematics")
        return o["com"], o["foot_lf"], o["foot_rf"]

    def integrate(self, q: torch.Tensor, v: torch.Tensor, dv: torch.Tensor, dt: float) -> None:
        """In place: v_mean = v + dt/2 dv; v += dt dv; q = q (+) dt v_mean (ref:ctrl/WalkController.py:291-295)."""
        n = q.shape[0]
        self._chk(q, n, self.nq, "q")
        self._chk(v, n, self.nv, "v")
        self._chk(dv, n, self.nv, "dv")
        check(self.lib.tsidb_integrate(self.h, n, 0, q.data_ptr(), v.data_ptr(), dv.data_ptr(), float(dt), self._stream()),
              "tsidb_integrate")

    # ------------------------------------------------------------------ gait phase machine / closed-loop rollout
    def gait_reset(self, n: int, dt: float, step_duration: float, step_length: float, step_height: float, com_height: float,
                   phase0: Optional[torch.Tensor] = None, vcmd: Optional[torch.Tensor] = None) -> None:
        """(Re)start the device gait of n envs from the default references (tsidb_gait_reset).  phase0 [n] in
        [0,1) and vcmd [n,2] are CUDA fp64 tensors (None = zeros)."""
        gc = TsidbGaitConf(float(dt), float(step_duration), float(step_length), float(step_height), float(com_height))
        if phase0 is not None:
            self._chk(phase0.reshape(n, 1), n, 1, "phase0")
        if vcmd is not None:
            self._chk(vcmd, n, 2, "vcmd")
        self._gait_n = n
        check(self.lib.tsidb_gait_reset(self.h, n, C.byref(gc), phase0.data_ptr() if phase0 is not None else None,
                                        vcmd.data_ptr() if vcmd is not None else None, self._stream()), "tsidb_gait_reset")

    def gait_set_plan(self, steps: Optional[torch.Tensor], n_steps: Optional[torch.Tensor] = None, rise_ratio: float = 0.5) -> None:
        """Footstep plan [N,S,4] (x, y, yaw, side) + n_steps [N] int32 (e.g. from footstep_plan) for the device gait: the
        swing foot follows FootTrajectory(rise_ratio) to the env's next footstep of its side.  None: straight steps."""
        if steps is None:
            check(self.lib.tsidb_gait_set_plan(self.h, 0, None, None, 0, 0.5, self._stream()), "tsidb_gait_set_plan")
            return
        n, S = steps.shape[0], steps.shape[1]
        self._chk(steps.reshape(n, 4 * S), n, 4 * S, "steps")
        if n_steps.dtype != torch.int32 or tuple(n_steps.shape) != (n,) or n_steps.device != self.device:
            raise TypeError("n_steps: expected an int32 [N] tensor on the engine's device")
        check(self.lib.tsidb_gait_set_plan(self.h, n, steps.data_ptr(), n_steps.data_ptr(), S, float(rise_ratio), self._stream()),
              "tsidb_gait_set_plan")

    def _view(self, ptr: int, shape, dtype: torch.dtype) -> torch.Tensor:
        """A torch view of a library-owned device array (no copy)."""
        np_dt = {torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]

        class _Arr:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": np_dt, "data": (int(ptr), False), "version": 2}

        return torch.as_tensor(_Arr(), device=self.device)

    def gait_state(self) -> Dict[str, torch.Tensor]:
        """Views of the gait's device state: references (com, foot_lf, foot_rf, contact_lf, contact_rf), mask,
        phase and the per-env count of failed ticks."""
        n = self._gait_n
        r = TsidbRefs()
        m, ph, fl = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self.lib.tsidb_gait_state(self.h, C.byref(r), C.byref(m), C.byref(ph), C.byref(fl)), "tsidb_gait_state")
        out = {k: self._view(getattr(r, k), (n, nd), torch.float64)
               for k, nd in (("com", 9), ("foot_lf", 24), ("foot_rf", 24), ("contact_lf", 12), ("contact_rf", 12))}
        out["mask"] = self._view(m.value, (n,), torch.uint8)
        out["phase"] = self._view(ph.value, (n,), torch.float64)
        out["fails"] = self._view(fl.value, (n,), torch.int32)
        return out

    def gait_step(self, foot_lf_now: torch.Tensor, foot_rf_now: torch.Tensor, status: Optional[torch.Tensor] = None) -> None:
        n = self._gait_n
        self._chk(foot_lf_now, n, 12, "foot_lf_now")
        self._chk(foot_rf_now, n, 12, "foot_rf_now")
        check(self.lib.tsidb_gait_step(self.h, n, foot_lf_now.data_ptr(), foot_rf_now.data_ptr(),
                                       status.data_ptr() if status is not None else None, self._stream()), "tsidb_gait_step")

    def foot_trajectory(self, t0: float, t1: float, start: torch.Tensor, target: torch.Tensor, step_height: float,
                        rise_ratio: float, t: torch.Tensor) -> torch.Tensor:
        """FootTrajectory([t0, t1], start, target, step_height, rise_ratio) of ref:ctrl/Foot_Trajectory.py:6-43 evaluated
        per env on the device: start/target [N,4] = (x, y, z, yaw), t [N]; returns [N,4,4]: value and the first three
        derivatives (axis 1) of (x, y, z, yaw) (axis 2)."""
        n = start.shape[0]
        self._chk(start, n, 4, "start")
        self._chk(target, n, 4, "target")
        self._chk(t.reshape(n, 1), n, 1, "t")
        out = torch.empty((n, 16), dtype=torch.float64, device=self.device)
        check(self.lib.tsidb_foot_trajectory(self.h, n, float(t0), float(t1), start.data_ptr(), target.data_ptr(), float(step_height),
                                             float(rise_ratio), t.data_ptr(), out.data_ptr(), self._stream()), "tsidb_foot_trajectory")
        return out.reshape(n, 4, 4)

    def footstep_plan(self, path: torch.Tensor, init_supports: torch.Tensor, step_length: float, step_width: float,
                      n_pts: Optional[torch.Tensor] = None, max_steps: int = 64) -> Tuple[torch.Tensor, torch.Tensor]:
        """FootstepPlanner(step_width, step_length).plan(path, init_supports) of ref:ctrl/Footstep_Planner.py:92-125 per env
        on the device: path [N,P,2], init_supports [N,2,4] = (x, y, yaw, side); returns (steps [N,max_steps,4],
        n_steps [N] int32)."""
        n, P = path.shape[0], path.shape[1]
        self._chk(path.reshape(n, 2 * P), n, 2 * P, "path")
        self._chk(init_supports.reshape(n, 8), n, 8, "init_supports")
        steps = torch.zeros((n, max_steps, 4), dtype=torch.float64, device=self.device)
        ns = torch.zeros(n, dtype=torch.int32, device=self.device)
        check(self.lib.tsidb_footstep_plan(self.h, n, path.data_ptr(), n_pts.data_ptr() if n_pts is not None else None, P,
                                           init_supports.data_ptr(), float(step_length), float(step_width), steps.data_ptr(),
                                           ns.data_ptr(), int(max_steps), self._stream()), "tsidb_footstep_plan")
        return steps, ns

    def rollout(self, q: torch.Tensor, v: torch.Tensor, n_steps: int, use_graph: bool = True) -> TickOutput:
        """n_steps closed-loop ticks on the device (tick -> integrate_dv -> gait step), q and v advanced in place;
        returns the last step's outputs (tau, ddq, f, status, iters; the engine's cached buffers, overwritten by the
        next call of the same batch size).  Needs gait_reset first."""
        n = q.shape[0]
        self._chk(q, n, self.nq, "q")
        self._chk(v, n, self.nv, "v")
        o = self._outputs(n, False)
        check(self.lib.tsidb_rollout(self.h, n, int(n_steps), q.data_ptr(), v.data_ptr(), o["tau"].data_ptr(), o["ddq"].data_ptr(),
                                     o["f"].data_ptr(), o["status"].data_ptr(), o["iters"].data_ptr(), int(use_graph),
                                     self._stream()), "tsidb_rollout")
        return self._tick_output(o, False, False)

    def diagnostics(self, out: TickOutput, contact_mask: Optional[torch.Tensor], omega: float) -> Dict[str, torch.Tensor]:
        """CoP [N,3], capture point [N,3] and support (lf.xy, rf.xy) [N,4] of a tick computed with aux=True
        (tsidb_diagnostics; ref:ctrl/WalkController.py:255-289, ref:legacy/biped.py:224-234)."""
        if out.wrench is None or out.com is None:
            raise RuntimeError("diagnostics needs a tick computed with aux outputs")
        n = out.com.shape[0]
        f64 = dict(dtype=torch.float64, device=self.device)
        res = {"cop": torch.empty((n, 3), **f64), "capture_point": torch.empty((n, 3), **f64), "support": torch.empty((n, 4), **f64)}
        a = TsidbAuxOut()
        a.com, a.foot_lf, a.foot_rf, a.wrench = out.com.data_ptr(), out.foot_lf.data_ptr(), out.foot_rf.data_ptr(), out.wrench.data_ptr()
        check(self.lib.tsidb_diagnostics(self.h, n, C.byref(a), contact_mask.data_ptr() if contact_mask is not None else None,
                                         float(omega), res["cop"].data_ptr(), res["capture_point"].data_ptr(),
                                         res["support"].data_ptr(), self._stream()), "tsidb_diagnostics")
        return res

    def ci_row(self, block: int, side: int, i: int) -> int:
        return int(self.lib.tsidb_ci_row(self.h, block, side, i))

    def debug_terms(self, env: int, n_contacts: int) -> Dict[str, np.ndarray]:
        """M, nle, JF (LF, RF sole Jacobians, LOCAL), the dv block of the Hessian and the dv part of the gradient that the
        dynamics kernel produced for `env` in the last tick (tsidb_debug_terms; parity tests)."""
        nv = self.nv
        out = {"M": np.empty((nv, nv)), "nle": np.empty(nv), "JF": np.empty((2, 6, nv)), "H": np.empty((nv, nv)), "g": np.empty(nv)}
        check(self.lib.tsidb_debug_terms(self.h, env, n_contacts, *(out[k].ctypes.data for k in ("M", "nle", "JF", "H", "g"))),
              "tsidb_debug_terms")
        return out

    def launch_count(self) -> int:
        return int(self.lib.tsidb_launch_count(self.h))

    KERNEL_NAMES = ("class_sort", "dynamics", "eliminate", "j2", "activeset")

    def set_sched_hint(self, on: bool) -> None:
        """Longest-first order of the envs inside a contact class by the previous tick's iteration counts (tsidb.h)."""
        check(self.lib.tsidb_set_sched_hint(self.h, int(on)), "tsidb_set_sched_hint")

    def set_timing(self, on: bool) -> None:
        check(self.lib.tsidb_set_timing(self.h, int(on)), "tsidb_set_timing")

    def last_tick_ms(self) -> Dict[str, float]:
        """CUDA-event durations of the kernels of the last tick (needs set_timing(True))."""
        ms = (C.c_float * 5)()
        check(self.lib.tsidb_last_tick_ms(self.h, ms), "tsidb_last_tick_ms")
        return {k: float(ms[i]) for i, k in enumerate(self.KERNEL_NAMES)}


def fp64_peak_tflops(device: int = 0) -> float:
    lib = load_library()
    out = C.c_double(0.0)
    check(lib.tsidb_fp64_peak(device, C.byref(out)), "tsidb_fp64_peak")
    return float(out.value)
