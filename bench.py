#!/usr/bin/env python
"""Headline benchmark: Fmax+LPT Mcells/s of the collapse-time hot path (BASELINE.json).

One step = one pass of the hot path over one synthetic box: the S smoothing radii of
compute_fmax (fused k-space kernel -> 3-D FFT -> collapse -> running max) followed by the
3LPT displacement stage, on the delta_k that GenIC left resident in HBM (GenIC is excluded
from the metric, SURVEY.md 8d, and reported separately).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                   # CPU arm: the oracle port on host cores

Prints ONE JSON line (see the task contract): value = device-timed throughput with inputs
resident; e2e = the same metric through the reference-facing call with host buffers (H2D of
kdensity from pinned memory + compute + D2H of the AoS products[]).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
METRIC = "Fmax+LPT Mcells/s"


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def default_grid(ngpus: int) -> int:
    """1024^3 fits one B200 (17 fields of 8.7 GB at peak); 2048^3 needs the 8 GPUs of the box
    (one FP64 field = 68.7 GB).  2 and 4 GPUs run the 1024^3 box slab-decomposed (strong scaling)."""
    return 2048 if ngpus >= 8 else 1024


def workload_config(N: int, world: int, S: int) -> dict:
    """`config` of the JSON line, shared by both arms (the reference arm times a bounded sample of it)."""
    lx = N // world
    return {"workload": f"synthetic {N}^3 grid, full smoothing-radius sweep (S={S}) + 3LPT on {world} B200"
                        + (f", x-slabs of {lx} planes, FFT transposes as peer stores over NVLink" if world > 1 else ""),
            "grid": N, "nsmooth": S, "lpt_order": 3, "cosmology": "HMF_Validation (EH, Omega0=.25, h=.7, sigma8=.8)",
            "seed": 486604, "parallelism": f"slab{world}" if world > 1 else "single GPU",
            "l2_policy": f"inputs larger than L2 (each field {8 * N ** 3 / world / 1e9:.1f} GB per GPU)"}


def algorithmic_bytes(N: int, S: int):
    """Per-kernel algorithmic HBM bytes of the radius loop and SURVEY 8d totals."""
    Nr = float(N) ** 3
    Nc = float(N) * N * (N // 2)          # half-complex elements the Hessian passes touch (kz < N/2)
    Ncs = float(N) * N * (N // 2 + 1)     # SURVEY.md 8(d) definition
    per = {
        "xpass_kernel": 16 * Nc + 3 * 16 * Nc,           # read delta_k once, write 3 x-transformed fields
        "ypass_kernel": 3 * 16 * Nc + 6 * 16 * Nc,       # read 3, write 6
        "zpass_collapse_kernel": 6 * 16 * Nc + 12 * Nr,  # read 6, Fmax r/w (8 B) + Rmax w (4 B)
    }
    survey_rad = 400 * Ncs + 12 * Nr
    survey_lpt = 1680 * Ncs + 288 * Nr
    return per, survey_rad, survey_lpt, S * survey_rad + survey_lpt


PARITY_FIXTURE = ROOT / "tests" / "golden" / "b200_1gpu_fmax_1024.json"


def parity_vs_fixture(N: int, world: int, tvar, pdf) -> dict:
    """TrueVariance and FmaxPDF of this run against the committed single-GPU run of the same box."""
    if not PARITY_FIXTURE.exists():
        return {"parity_vs_1gpu": None, "parity_note": "no fixture (tests/golden/b200_1gpu_fmax_1024.json)"}
    fx = json.loads(PARITY_FIXTURE.read_text())
    if fx["grid"] != N:
        return {"parity_vs_1gpu": None, "parity_note": f"fixture is a {fx['grid']}^3 run, this one {N}^3"}
    tv0, pdf0 = np.array(fx["true_variance"]), np.array(fx["fmax_pdf"], dtype=np.int64)
    dv = float(np.abs(np.asarray(tvar) / tv0 - 1.0).max())
    dp = int(np.abs(np.asarray(pdf, dtype=np.int64) - pdf0).max())
    return {"parity_vs_1gpu": bool(dv <= 1e-10 and dp <= 2), "parity_true_variance_max_rel": dv, "parity_fmaxpdf_max_bin_diff": dp,
            "parity_fixture_gpus": fx["n_gpus"]}


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pinocchio_b200.build import build
    build()
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import PRODUCT_DTYPE_3LPT, Pinocchio, RunConfig

    N = args.grid or default_grid(world)
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    cfg = RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3)
    lad = SmoothingLadder(np.array(HMF_RADII), np.zeros(len(HMF_RADII)))
    pin = Pinocchio(cfg, cosmo, device=local, smoothing=lad, rank=rank, nranks=world)
    lx = N // world
    stream = torch.cuda.current_stream()
    pin.set_stream(stream.cuda_stream)
    S = lad.Nsmooth

    t0 = time.time()
    pin.GenIC_large()
    genic_s = pin.timers().dens

    def step():
        pin.compute_fmax(displacements=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    tm0 = pin.timers()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    tm1 = pin.timers()
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    cells = float(N) ** 3              # one box, slab-decomposed over the ranks
    value = cells / (ms_per_step * 1e-3) / 1e6

    # ---- size-independent sanity of the timed result (not timed; Fmax_PDF all-reduces over the ranks)
    pdf = pin.Fmax_PDF()
    tvar = np.asarray(pin.TrueVariance, dtype=np.float64)
    checks = {"pdf_total_is_ncells": bool(int(pdf.sum()) == N ** 3),
              "variance_ladder_monotone": bool(np.isfinite(tvar).all() and (np.diff(tvar) > 0).all()),
              "collapsed_fraction": round(float(pdf[10:].sum()) / float(N) ** 3, 6)}
    # the measured variance of the smoothed field against the one the host cosmology predicts for each radius
    # (the reference logs both as "expected sigma" / "computed sigma", src/fmax.c:141-147): a size-independent
    # check of the whole FFT path that also holds for the 2048^3 box no single GPU can cross-check
    try:
        from pinocchio_b200.cosmology import set_smoothing
        theo = set_smoothing(cosmo, 1.0 / 0.7)
        if theo.Variance.size == tvar.size:
            big = theo.Radius >= 2.0 / 0.7      # radii of at least two cells: below, the grid's own cut-off in k dominates
            checks["sigma_vs_theory_max_rel"] = round(float(np.abs(np.sqrt(tvar / theo.Variance) - 1.0)[big].max()), 5)
            checks["sigma_vs_theory_radii"] = int(big.sum())
    except Exception as ex:  # noqa: BLE001
        checks["sigma_vs_theory_max_rel"] = str(ex)[:100]
    # parity of the slab-decomposed runs with the single-GPU run of the same 1024^3 box (fixture written by
    # `bench.py --write-parity-fixture` on one GPU): TrueVariance[S] and the 210-bin Fmax histogram
    if rank == 0:
        checks.update(parity_vs_fixture(N, world, tvar, pdf))
        if args.write_parity_fixture and world == 1:
            Path(args.write_parity_fixture).write_text(json.dumps({"grid": N, "n_gpus": 1, "radii": HMF_RADII, "seed": 486604,
                                                                   "true_variance": [float(v) for v in tvar],
                                                                   "fmax_pdf": [int(v) for v in pdf]}))

    # ---- per-kernel device times measured live (CUDA events inside the engine, same stream)
    K = args.steps
    per_launch_ms = {"xpass_kernel": (tm1.hess_x - tm0.hess_x) / (K * S) * 1e3,
                     "ypass_kernel": (tm1.hess_y - tm0.hess_y) / (K * S) * 1e3,
                     "zpass_collapse_kernel": (tm1.hess_z - tm0.hess_z) / (K * S) * 1e3}
    abytes, survey_rad, survey_lpt, survey_total = algorithmic_bytes(N, S)
    abytes = {k: v / world for k, v in abytes.items()}            # per rank
    survey_rad, survey_lpt, survey_total = survey_rad / world, survey_lpt / world, survey_total / world
    peak, peak_src = measured_peaks()
    dom = max(per_launch_ms, key=per_launch_ms.get)
    achieved = abytes[dom] / (per_launch_ms[dom] * 1e-3) / 1e9
    rad_ms = sum(per_launch_ms.values())
    lpt_ms = (tm1.lpt - tm0.lpt) / K * 1e3
    # DRAM bytes and instruction counts per launch from the committed ncu counter pass of this build
    # (profiles/r02_traffic.json, written by tools/ncu_to_traffic.py; same grid / GPU count only)
    traffic, traffic_src, prof = None, None, {}
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            prof = json.loads((ROOT / "profiles" / name).read_text())
            tj = prof.get(dom)
            if tj and tj["grid"] == N and tj["n_gpus"] == world:
                traffic, traffic_src = round((tj["read_gb"] + tj["write_gb"]) * 1e9), f"profiles/{name}: {tj['source']}"
            break
        except (OSError, ValueError, KeyError):
            continue
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": round(abytes[dom]), "peak_source": peak_src,
                "note": "the dominant kernel (z pass + collapse epilogue) is bound by instruction issue, not by HBM: an FP64 "
                        "instruction holds a sub-partition's dispatch port for two cycles on sm_100 (tools/fp64lat.cu), and "
                        "2 x FP64 + other instructions per cell account for its whole run time (roofline.issue); frac is "
                        "against the HBM peak as the contract defines it",
                "ms_per_launch": {k: round(v, 3) for k, v in per_launch_ms.items()},
                "gbs_per_kernel": {k: round(abytes[k] / (v * 1e-3) / 1e9, 1) for k, v in per_launch_ms.items()},
                "per_radius_ms": round(rad_ms, 3),
                "per_radius_frac_of_survey_roofline": round(survey_rad / (rad_ms * 1e-3) / 1e9 / peak, 4),
                "lpt_stage_ms": round(lpt_ms, 3),
                # the engine's own split of the displacement stage: sources + k-vectors (three r2c, the contraction),
                # the four velocity triples, and inside those the four inverse x passes (peer stores at N > 1)
                "lpt_breakdown_ms": {"sources_and_kvectors": round((tm1.disp_sources - tm0.disp_sources) / K * 1e3, 2),
                                     "velocities": round((tm1.disp_vel - tm0.disp_vel) / K * 1e3, 2),
                                     "velocity_x_passes": round((tm1.disp_x - tm0.disp_x) / K * 1e3, 2)},
                "lpt_frac_of_survey_roofline": round(survey_lpt / (lpt_ms * 1e-3) / 1e9 / peak, 4),
                "whole_step_frac_of_survey_roofline": round(survey_total / (ms_per_step * 1e-3) / 1e9 / peak, 4)}
    # Instruction-issue roofline of every radius kernel: a sub-partition dispatches one warp instruction per clock and an
    # FP64 one occupies the port for two (measured: 0.5 warp-DFMA / clk / sub-partition, 8-cycle dependent latency), so
    # a kernel needs at least (2 x FP64 + other) warp instructions / (592 sub-partitions x clock).  Instruction counts:
    # ncu counter pass of this build (smsp__inst_executed.sum, sm__inst_executed_pipe_fp64.sum), time: live.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue = {}
    for k, ms_k in per_launch_ms.items():
        e = prof.get(k)
        if not e or not e.get("warp_inst") or e.get("grid") != N:
            continue
        scale = 1.0 / world if e.get("n_gpus", 1) == 1 else 1.0        # counts of the 1-GPU launch cover the whole box
        slots = (e["warp_inst"] + e["fp64_warp_inst"]) * scale
        floor_ms = slots / (148 * 4 * sm_mhz * 1e6) * 1e3
        issue[k] = {"warp_inst": int(e["warp_inst"] * scale), "fp64_warp_inst": int(e["fp64_warp_inst"] * scale),
                    "issue_floor_ms": round(floor_ms, 2), "frac_of_issue_roofline": round(floor_ms / ms_k, 4)}
        if "fp64_inst_per_cell" in e:
            issue[k]["fp64_inst_per_cell"] = round(e["fp64_inst_per_cell"], 1)
            issue[k]["inst_per_cell"] = round(e["inst_per_cell"], 1)
    roofline["issue"] = dict(issue, sm_mhz=sm_mhz, source=prof.get("_source"),
                             model="ms >= (warp_inst + fp64_warp_inst) / (592 x sm_clock): FP64 instructions take two dispatch cycles")
    if world > 1:
        # NVLink side of the metric: the only exchange is the transpose fused into the x pass as
        # peer stores.  Per GPU and radius this design moves 3 x-transformed fields (the y pass makes the 6
        # Hessian components from them), i.e. 3*16*Nc*(P-1)/P^2 bytes out and as many in;
        # SURVEY's unfused model counts one transpose per FFT = 6 fields.  `achieved` divides the bytes a GPU
        # sends by the whole x-pass kernel time (line FFTs and local stores included), so it is a lower bound
        # on the link rate.  Combined roofline of a radius = max(HBM time, NVLink time), perfect overlap.
        nvl_peak = 900.0
        Ncs = float(N) * N * (N // 2 + 1)
        sent = 3 * 16 * Ncs * (world - 1) / world ** 2
        sent_survey = 6 * 16 * Ncs * (world - 1) / world ** 2
        t_hbm = survey_rad / (peak * 1e9)
        t_nvl, t_nvl_survey = sent / (nvl_peak * 1e9), sent_survey / (nvl_peak * 1e9)
        xfer_ms = (tm1.xfer - tm0.xfer) / (K * S) * 1e3                    # copy-engine transposes per radius (0: peer-store schedule)
        roofline["nvlink"] = {"kernel": "copy-engine transposes behind a local x pass, hidden under the collapse pass" if xfer_ms > 0
                              else "xpass_kernel (transpose fused as peer stores)", "unit": "GB/s", "peak": nvl_peak,
                              "transposes_ms_per_radius": round(xfer_ms, 3),
                              "transposes_gbs": round(sent / (xfer_ms * 1e-3) / 1e9, 1) if xfer_ms > 0 else None,
                              "peak_source": "nominal NVLink 5, per direction per GPU",
                              "bytes_sent_per_gpu_per_radius": round(sent),
                              "bytes_sent_per_gpu_per_radius_survey_model": round(sent_survey),
                              "achieved": round(sent / (per_launch_ms["xpass_kernel"] * 1e-3) / 1e9, 1),
                              "frac": round(sent / (per_launch_ms["xpass_kernel"] * 1e-3) / 1e9 / nvl_peak, 4),
                              "per_radius_frac_of_combined_roofline": round(max(t_hbm, t_nvl) / (rad_ms * 1e-3), 4),
                              "per_radius_frac_of_combined_roofline_survey_model": round(max(t_hbm, t_nvl_survey) / (rad_ms * 1e-3), 4),
                              "per_radius_frac_of_summed_roofline": round((t_hbm + t_nvl) / (rad_ms * 1e-3), 4)}
    launches = int(tm1.kernel_launches - tm0.kernel_launches)

    # ---- e2e: host buffers through the reference-facing calls.  One step = H2D of kdensity from pinned memory,
    #      the sweep + 3LPT, and the hand-off to the fragmentation as the drop-in does it
    #      (shim/fmax_b200.c download_products_compact): D2H of Fmax of every cell, of the cell indices of the
    #      collapsed cells (Fmax >= Flast = 1, selected and ordered on the device) and of their 56-byte records --
    #      what src/distribute.c reads of products[].  Fmax and the index list are final before the displacement
    #      stage starts, so their selection, sort and downloads are started there (pinb200_handoff_begin) and hide
    #      under it; the records follow.  The plain copy of all 56-byte records (r01's e2e) is timed once beside
    #      it (`full_aos`).
    e2e = None
    if not args.no_e2e:
        import ctypes
        from pinocchio_b200.engine import _PD, ProductLayout
        ncell_local = lx * N * N
        f = PRODUCT_DTYPE_3LPT.fields
        lay = ProductLayout(56, 4, f["Rmax"][1], f["Fmax"][1], f["Vel"][1], f["Vel_2LPT"][1], f["Vel_3LPT_1"][1],
                            f["Vel_3LPT_2"][1])
        flast = 1.0
        cnt = ctypes.c_size_t(0)
        pin._ck(pin.lib.pinb200_collapsed_cells(pin.h, flast, None, 0, ctypes.byref(cnt)))     # same field every step
        ncoll = int(cnt.value)
        err = ""
        try:   # pinned host buffers: kdensity slab (8.6 GB), Fmax (4.3 GB), indices + records of the collapsed cells (38 GB)
            kd_host = torch.empty((N, lx, N // 2 + 1, 2), dtype=torch.float64, pin_memory=True)
            fm_host = torch.empty((ncell_local,), dtype=torch.float32, pin_memory=True)
            idx_host = torch.empty((max(ncoll, 1),), dtype=torch.int32, pin_memory=True)
            rec_host = torch.empty((max(ncoll, 1) * 56,), dtype=torch.uint8, pin_memory=True)
        except Exception as ex:
            err = str(ex)[:200]
        okf = torch.tensor([0 if err else 1], device="cuda")
        if world > 1:
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if int(okf.item()) == 0:   # every rank skips together: the compute calls contain cross-GPU barriers
            e2e = {"value": None, "unit": "Mcells/s", "error": err or "pinned host allocation failed on a peer rank"}
        else:
            pin._ck(pin.lib.pinb200_download_kdensity(pin.h, ctypes.cast(kd_host.data_ptr(), _PD)))
            rec_np = rec_host.numpy().view(PRODUCT_DTYPE_3LPT)
            PU = ctypes.POINTER(ctypes.c_uint)

            phases = {"h2d_kdensity": 0.0, "compute_with_handoff_prefetch_under_lpt": 0.0, "handoff_wait": 0.0, "d2h_records": 0.0}
            tme0 = pin.timers()

            def e2e_step():
                t = [time.perf_counter()]          # every library call returns with its work done: wall-clock phases
                pin._ck(pin.lib.pinb200_upload_kdensity(pin.h, ctypes.cast(kd_host.data_ptr(), _PD)))
                t.append(time.perf_counter())
                pin.compute_fmax(displacements=False)
                # Fmax is final: its download, the selection + sort and the download of the index list run on side
                # streams / the copy engine UNDER the displacement stage (pinb200_handoff_begin .. _end)
                c = ctypes.c_size_t(0)
                pin._ck(pin.lib.pinb200_handoff_begin(pin.h, flast, ctypes.cast(fm_host.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                                      ctypes.cast(idx_host.data_ptr(), PU), ncoll, ctypes.byref(c)))
                assert int(c.value) == ncoll
                pin.compute_displacements(1, 0, pin.cfg.segment_redshift)
                t.append(time.perf_counter())
                pin._ck(pin.lib.pinb200_handoff_end(pin.h))
                t.append(time.perf_counter())
                pin._ck(pin.lib.pinb200_download_products_sorted(pin.h, ctypes.c_void_p(rec_host.data_ptr()), ctypes.byref(lay), 0, ncoll))
                t.append(time.perf_counter())
                for k, name in enumerate(phases):
                    phases[name] += t[k + 1] - t[k]
                return float(rec_np["Fmax"][0]) + float(fm_host[0])

            # A failing library call (an allocation, say) fails on every rank alike -- same sizes, same order of calls --
            # so all ranks leave the block together and the line is still printed, with the reason in e2e.error.
            e2e_err = ""
            ne = max(1, min(args.steps, 3))
            w = float("nan")
            try:
                e2e_step()
                barrier()
                for k in phases:
                    phases[k] = 0.0
                w0 = time.perf_counter()
                for _ in range(ne):
                    e2e_step()
                barrier()
                w = (time.perf_counter() - w0) / ne
            except Exception as ex:  # noqa: BLE001
                e2e_err = str(ex)[:300] or type(ex).__name__
            okf = torch.tensor([0 if e2e_err else 1], device="cuda")
            if world > 1:
                dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if int(okf.item()) == 0:
                e2e = {"value": None, "unit": "Mcells/s", "error": e2e_err or "the e2e step failed on a peer rank"}
                del kd_host, fm_host, idx_host, rec_host
            if e2e is None:
                sort_ms = pin.timers().sort_ms
                tw = torch.tensor([w], device="cuda", dtype=torch.float64)
                tc = torch.tensor([float(ncoll)], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
                    dist.all_reduce(tc, op=dist.ReduceOp.SUM)
                ncoll_all = int(tc.item())
                # the hand-off is what the fragmentation would get: ordered, all collapsed cells, records of the right cells
                fs = rec_np["Fmax"][:ncoll]
                order_ok = bool(ncoll == 0 or ((np.diff(fs[:: max(1, ncoll // (1 << 22))]) <= 0).all() and fs[-1] >= flast
                                               and np.array_equal(fs[:4096], fm_host.numpy()[idx_host.numpy()[:4096].view(np.uint32)])))
                e2e = {"value": round(cells / float(tw.item()) / 1e6, 2), "unit": "Mcells/s",
                       "h2d_bytes_per_step": int(N * N * (N // 2 + 1) * 16),
                       "d2h_bytes_per_step": int(4 * N ** 3 + 60 * ncoll_all),
                       "steps": ne, "ms_per_step": round(float(tw.item()) * 1e3, 2),
                       "handoff": "Fmax of every cell + index list and 56-byte records of the cells with Fmax >= 1 in order of descending "
                                  "Fmax (device-side selection + radix sort), as shim/fmax_b200.c fills products[] for src/distribute.c",
                       "collapsed_cells": ncoll_all, "select_sort_ms_device": round(float(sort_ms), 2), "handoff_ok": order_ok,
                       "phases_ms_rank0": {k: round(v / ne * 1e3, 1) for k, v in phases.items()}}
                tme1 = pin.timers()
                # the compute phase on the device clock (CUDA events inside the engine) next to its wall-clock time above
                e2e["phases_ms_rank0"]["compute_device_events"] = round(((tme1.fmax - tme0.fmax) + (tme1.lpt - tme0.lpt)) * 1e3 / (ne + 1), 1)
                # the plain copy of every record, once, for comparison (r01's e2e definition)
                try:
                    chunk = min(ncell_local, 1 << 26)
                    nst = min(chunk * 56, rec_host.numel())
                    chunk = nst // 56
                    barrier()
                    w0 = time.perf_counter()
                    pin._ck(pin.lib.pinb200_upload_kdensity(pin.h, ctypes.cast(kd_host.data_ptr(), _PD)))
                    pin.compute_fmax(displacements=True)
                    for b0 in range(0, ncell_local, chunk):
                        nn = min(chunk, ncell_local - b0)
                        pin._ck(pin.lib.pinb200_download_products(pin.h, ctypes.c_void_p(rec_host.data_ptr()), ctypes.byref(lay), b0, nn))
                    barrier()
                    wf = time.perf_counter() - w0
                    twf = torch.tensor([wf], device="cuda", dtype=torch.float64)
                    if world > 1:
                        dist.all_reduce(twf, op=dist.ReduceOp.MAX)
                    e2e["full_aos"] = {"ms_per_step": round(float(twf.item()) * 1e3, 2), "value": round(cells / float(twf.item()) / 1e6, 2),
                                       "d2h_bytes_per_step": int(N ** 3 * 56), "steps": 1}
                except Exception as ex:  # noqa: BLE001
                    e2e["full_aos"] = {"error": str(ex)[:200]}
                # the e2e step is PCIe-bound: report the bare pinned-copy rates of this box beside it, so that
                # (h2d_bytes / h2d_gbs + d2h_bytes / d2h_gbs + device step) can be compared with ms_per_step
                try:
                    nb = int(min(rec_host.numel(), 1 << 30))
                    dbuf = torch.empty(nb, dtype=torch.uint8, device="cuda")
                    rates = {}
                    for name, (dst, src) in {"d2h_gbs": (rec_host[:nb], dbuf), "h2d_gbs": (dbuf, rec_host[:nb])}.items():
                        dst.copy_(src, non_blocking=True)
                        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ev0.record()
                        for _ in range(3):
                            dst.copy_(src, non_blocking=True)
                        ev1.record()
                        ev1.synchronize()
                        rates[name] = round(3 * nb / (ev0.elapsed_time(ev1) * 1e-3) / 1e9, 1)
                    del dbuf
                    e2e["pcie_pinned_copy"] = dict(rates, bytes=nb)
                    pcie_ms = (e2e["h2d_bytes_per_step"] / rates["h2d_gbs"] + e2e["d2h_bytes_per_step"] / rates["d2h_gbs"]) / world / 1e6
                    e2e["pcie_floor_ms_per_step"] = round(pcie_ms, 1)
                except Exception as ex:  # noqa: BLE001
                    e2e["pcie_pinned_copy"] = {"error": str(ex)[:200]}
                del kd_host, fm_host, idx_host, rec_host

    # ---- BASELINE.json configs[4]: scale-dependent growth (massive neutrinos / modified gravity) at the same grid:
    #      the displacement fields are re-derived per redshift segment with a growth rate G(|k|) evaluated per mode in
    #      the x-pass loader (xpass_growthk_kernel: InterpolateGrowth's ten k-bin splines, src/cosmo.c:1728-1757, a
    #      log10 and a pow per mode).  Timed next to the scale-independent velocity stage on the resident k-vectors.
    scaledep = None
    if world == 1 and not args.no_scaledep:
        try:
            nk = 10                                                        # NkBINS, LOGKMIN = -3, DELTALOGK = 0.5 (src/def_splines.h:40-42)
            g0 = np.log10(np.abs(pin.growth_rates(0.0)))
            tab = np.ascontiguousarray(g0[:, None] + 0.004 * np.arange(nk)[None, :])     # a few per cent of k dependence
            Nc = float(N) * N * (N // 2)
            xbytes = 3 * 16 * Nc                                           # one field read, two written (kx^0 and kx^1 jobs)
            res = {}
            for name, fn in (("scale_independent", None), ("scale_dependent", lambda z: tab)):
                pin.set_scale_dependent_growth(fn, -3.0, 0.5)
                pin.compute_displacements(0, 0, 0.0)                      # warm-up
                ta = pin.timers()
                for _ in range(2):
                    pin.compute_displacements(0, 0, 0.0)
                tb = pin.timers()
                xms = (tb.disp_x - ta.disp_x) * 1e3 / 8.0                 # four x passes per call
                res[name] = {"velocity_stage_ms": round((tb.disp_vel - ta.disp_vel) * 1e3 / 2.0, 2),
                             "xpass_ms_per_launch": round(xms, 3), "xpass_gbs": round(xbytes / (xms * 1e-3) / 1e9, 1)}
            pin.set_scale_dependent_growth(None)
            scaledep = {"workload": f"scale-dependent growth {N}^3: four first-derivative triples (Zel'dovich, 2LPT, 3LPT_1, 3LPT_2) "
                                    "with G(|k|) per mode, NkBINS = 10 tables as the reference hands them over per redshift segment",
                        "kernel": "xpass_growthk_kernel", "algorithmic_bytes_per_xpass": int(xbytes), **res,
                        "xpass_frac_of_hbm_peak": round(res["scale_dependent"]["xpass_gbs"] / peak, 4)}
            # the start-up side of the same configuration: set_scaledep_GM's 3 x Nsmooth x 210 variance integrals in
            # one device call (pinb200_scaledep_variances), synthetic tables of the shipped shapes
            from pinocchio_b200.engine import gauss_legendre_nodes, scaledep_variances
            nt, ns = 210, S
            tt = np.linspace(-2.2, 0.05, nt)
            lg = (1.0 - 0.05 * np.tanh(np.arange(nk) - 4.0)[:, None]) * tt[None, :]
            fo = 0.5 + 0.45 * (1.0 - 10.0 ** tt)[None, :] + 0.03 * np.tanh(np.arange(nk) - 5.0)[:, None]
            lk, w = gauss_legendre_nodes(-4.0, math.pi, 512, breaks=-3.0 + 0.5 * np.arange(nk))
            kk = 10.0 ** lk
            pk = np.array([cosmo.PowerSpectrum(float(v)) for v in kk[::8]])
            pk = np.exp(np.interp(lk, lk[::8], np.log(pk)))
            ad, ap = w * pk * kk ** 3 / (2 * math.pi ** 2), w * pk * kk / (2 * math.pi ** 2)
            rd, rp = np.array(HMF_RADII[:ns]), np.linspace(30.0, 0.0, ns)
            scaledep_variances(lk, ad, ap, lg, fo, -3.0, 0.5, rd, rp)      # warm-up (module load, allocations)
            t_sd = time.perf_counter()
            sv = scaledep_variances(lk, ad, ap, lg, fo, -3.0, 0.5, rd, rp)
            t_sd = time.perf_counter() - t_sd
            scaledep["startup_integrals"] = {"call": "pinb200_scaledep_variances", "integrals": int(3 * ns * nt), "nodes": int(lk.size),
                                             "wall_ms_incl_copies": round(t_sd * 1e3, 3), "finite": bool(np.isfinite(sv).all()),
                                             "replaces": "3 x Nsmooth x NBINS gsl_integration_qags calls of set_scaledep_GM "
                                                         "(src/initialization.c:1594-1601, 1742-1748, 1886-1892)"}
        except Exception as ex:  # noqa: BLE001
            scaledep = {"error": str(ex)[:300]} if scaledep is None else {**scaledep, "startup_integrals_error": str(ex)[:300]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_reference_sample(args.cpu_grid or 256, 1)

    # The probes below run code that round 1 could not time on hardware; their budget is bounded (120 + 180
    # + 150 s at worst) and the measurements above are parked on disk first, so a probe that misbehaves
    # cannot take the headline numbers with it.
    if rank == 0:
        try:
            (ROOT / "gpurun_out").mkdir(exist_ok=True)
            (ROOT / "gpurun_out" / "bench_before_probes.json").write_text(json.dumps(
                {"value": round(value, 2), "ms_per_step": round(ms_per_step, 3), "n_gpus": world, "grid": N, "e2e": e2e,
                 "roofline": roofline, "clocks": clocks, "cpu_baseline": cpu_baseline, "checks": checks}, default=float))
        except (OSError, TypeError, ValueError):
            pass

    # ---- device-side fragmentation hand-off, in a fresh process after this one
    #      has released the GPU: a failure there cannot touch the numbers above
    handoff = None
    if rank == 0 and world == 1 and not args.no_handoff:
        pin.close()
        torch.cuda.empty_cache()
        try:
            r = subprocess.run([sys.executable, str(ROOT / "scripts" / "gpu_handoff_probe.py"), str(N)], capture_output=True,
                               text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            handoff = json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            handoff = {"error": str(ex)[:300]}

    # ---- the linked drop-in against the reference program (whole PINOCCHIO run at 128^3), also isolated
    dropin = None
    if rank == 0 and world == 1 and not args.no_handoff:
        try:
            r = subprocess.run([sys.executable, str(ROOT / "scripts" / "dropin_probe.py")], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            dropin = json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            dropin = {"error": str(ex)[:300]}

    # ---- collapse-time tables (-DTABULATED_CT / -DELL_SNG): table build + tabulated sweep, isolated
    ctable = None
    if rank == 0 and world == 1 and not args.no_handoff:
        try:
            r = subprocess.run([sys.executable, str(ROOT / "scripts" / "gpu_ctable_probe.py"), str(N)], capture_output=True,
                               text=True, timeout=150)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            ctable = json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            ctable = {"error": str(ex)[:300]}

    # ---- N > 1 on a box other than the fixture's (2048^3 on 8 GPUs): the same ranks also run the 1024^3 box of the
    #      fixture once (sweep only, not timed), so that the slab decomposition is cross-checked against the
    #      single-GPU result in every scaling run
    if world > 1 and N != 1024 and PARITY_FIXTURE.exists():
        try:
            pin.close()
            torch.cuda.empty_cache()
            sub = Pinocchio(RunConfig(GridSize=1024, BoxSize_htrue=1024 / 0.7, lpt_order=3), cosmo, device=local, smoothing=lad,
                            rank=rank, nranks=world)
            sub.set_stream(stream.cuda_stream)
            sub.GenIC_large()
            sub.compute_fmax(displacements=False)
            spdf = sub.Fmax_PDF()
            stv = np.asarray(sub.TrueVariance, dtype=np.float64)
            sub.close()
            if rank == 0:
                checks["subrun_1024"] = parity_vs_fixture(1024, world, stv, spdf)
        except Exception as ex:  # noqa: BLE001
            checks["subrun_1024"] = {"error": str(ex)[:200]}

    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 2), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
               "scaling": "weak" if world in (1, 8) else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": workload_config(N, world, S),
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
               "cpu_baseline": cpu_baseline, "fragment_handoff": handoff, "dropin_program": dropin, "collapse_tables": ctable, "scaledep": scaledep, "checks": checks, "genic_s": round(genic_s, 4), "setup_s": round(time.time() - t0, 1)}
        print(json.dumps(out))
    pin.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_reference_sample(grid: int, steps: int, warmup: int = 0):
    """CPU baseline on a bounded sample of the same workload (full S-radius sweep + 3LPT on a
    `grid`^3 box), on all host cores.

    kind "reference": oracle/_ref, i.e. the reference's own src/{fmax,fmax-pfft,LPT,collapse_times}.c
    compiled verbatim (gcc -O3 -fopenmp) over one-task stand-ins for MPI/PFFT/GSL -- the FFT
    underneath is the in-repo Stockham of oracle/ref_fft.c, not FFTW.  Run in a fresh process
    (oracle/reference_runner.py) with OMP_NUM_THREADS = host cores.
    kind "port": the NumPy oracle, when oracle/_ref was never built."""
    N = grid
    cores = os.cpu_count() or 1
    runner = ROOT / "oracle" / "reference_runner.py"
    if (ROOT / "oracle" / "_ref" / "libpinocchio_ref.so").exists():
        env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="false")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, str(runner), "--grid", str(N), "--steps", str(steps), "--warmup", str(warmup),
                            "--threads", str(cores)], capture_output=True, text=True, env=env, timeout=3000)
        if r.returncode == 0:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            t = j["seconds_per_step"]
            return {"value": round(N ** 3 / t / 1e6, 4), "unit": "Mcells/s", "cores": cores, "kind": "reference",
                    "sample": f"{N}^3 box, same 9-radius sweep + 3LPT, the reference's own C sources (oracle/_ref: gcc -O3 "
                              f"-fopenmp, OpenMP on {cores} threads, one task; MPI/PFFT/GSL replaced by in-repo stand-ins, "
                              f"FFT = oracle/ref_fft.c not FFTW), {steps} step(s), {t:.2f} s/step",
                    "ms_per_step": round(t * 1e3, 1), "reference_timers_s": j["timers_last_step"]}
        sys.stderr.write(f"oracle/_ref run failed, falling back to the NumPy port:\n{r.stderr[-2000:]}\n")
    from oracle import pinocchio_oracle as po
    from pinocchio_b200.cosmology import Cosmology
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    kd = po.genic(N, N / 0.7, 486604, cosmo.PowerSpectrum)
    g = (cosmo.GrowingMode(0.0), cosmo.GrowingMode_2LPT(0.0), cosmo.GrowingMode_3LPT_1(0.0), cosmo.GrowingMode_3LPT_2(0.0))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        po.compute_fmax(kd, HMF_RADII, 1.0 / 0.7, cosmo.InverseGrowingMode, growth=g, lpt_order=3)
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    return {"value": round(N ** 3 / t / 1e6, 4), "unit": "Mcells/s", "cores": 1, "kind": "port",
            "sample": f"{N}^3 box, same 9-radius sweep + 3LPT, NumPy oracle (oracle/pinocchio_oracle.py), "
                      f"{steps} step(s), {t:.1f} s/step", "ms_per_step": round(t * 1e3, 1),
            "host_cores_available": os.cpu_count()}


def reference_sample_grid(steps: int, warmup: int, budget_s: float = 240.0) -> tuple[int, float]:
    """Grid of the bounded sample the reference arm times: the largest of 256^3 / 128^3 whose `steps + warmup`
    passes fit `budget_s` on this box's cores, estimated from one calibration pass at 128^3 (N^3 log N scaling)."""
    cal = cpu_reference_sample(128, 1)
    t128 = cal["ms_per_step"] * 1e-3
    t256 = t128 * 8.0 * 8.0 / 7.0
    return (256 if (steps + warmup) * t256 <= budget_s else 128), t128


def run_reference(args):
    """CPU arm: the reference's own C sources (oracle/_ref) on the host cores.  It runs EXACTLY `--steps` timed and
    `--warmup` untimed passes, each pass one bounded sample (a 256^3 or 128^3 box, chosen so that the whole run ends
    within a few minutes) of the B200 arm's workload, and says so in `config`: Mcells/s is size-normalised, but this
    is a sample, not the 1024^3 box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.cpu_grid > 0:
        N, t128 = args.cpu_grid, None
    else:
        N, t128 = reference_sample_grid(steps, warmup)
    cb = cpu_reference_sample(N, steps, warmup=warmup)
    S = len(HMF_RADII)
    cfg = workload_config(args.grid or default_grid(world), world, S)
    cfg["sample_of"] = cfg["workload"]
    cfg["workload"] = (f"bounded CPU sample of the B200 arm's workload: synthetic {N}^3 grid per step, the same smoothing-radius "
                       f"sweep (S={S}) + 3LPT, same cosmology and seed, one task on {cb['cores']} host threads")
    cfg["sample_grid"] = N
    if t128 is not None:
        cfg["sample_choice"] = f"calibration pass at 128^3: {t128:.2f} s; largest of 256^3/128^3 whose {steps}+{warmup} passes fit 240 s"
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Mcells/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": cfg, "cpu_baseline": cb, "gpu_launches": 0,
           "e2e": {"value": cb["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="override the grid side (default 1024)")
    ap.add_argument("--cpu-grid", type=int, default=0,
                    help="grid of the bounded CPU sample (0: 256 for the cpu_baseline leg; the reference arm picks 256 or 128 so that "
                         "steps + warmup passes fit a few minutes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-handoff", action="store_true", help="skip the fragmentation hand-off probe (fresh process, N=1 only)")
    ap.add_argument("--no-scaledep", action="store_true", help="skip the scale-dependent growth timing (configs[4], N=1 only)")
    ap.add_argument("--write-parity-fixture", default="", help="N=1: write TrueVariance and FmaxPDF of this run as the fixture the N>1 runs are compared with")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
