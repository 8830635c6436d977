#!/usr/bin/env python
"""
bench.py -- map_cells_to_space iterations/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5] [--precision bf16|fp32]
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference Mapper)

A "step" is one optimizer iteration (loss, backward, Adam) of the hot path on synthetic
expression-like inputs.  N>1: launched by torchrun, one rank per GPU, the
cells axis sharded (strong scaling: the total problem is fixed), one NCCL all-reduce per step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells, voxels, genes, types, clusters-mode, description)
    "c1": (26_431, 9_852, 249, 0, False,
           "reference fixtures data/test_ad_sc.h5ad x data/test_ad_sp.h5ad (tests/golden/c1_reference.npz), mode=cells"),
    "c2": (10_000, 1_000, 1_000, 0, False, "synthetic 10k cells x 1k voxels x 1k genes, mode=cells"),
    "c3": (100_000, 10_000, 2_000, 0, False, "synthetic 100k cells x 10k voxels x 2k genes, mode=cells"),
    "c4": (256, 50_000, 5_000, 0, True, "synthetic 256 clusters x 50k voxels x 5k genes, mode=clusters"),
    "c5": (50_000, 5_000, 2_000, 32, False,
           "synthetic 50k cells x 5k voxels x 2k genes, neighbourhood + ct-islands + Getis-Ord on"),
}
C5_LAMBDAS = dict(lambda_neighborhood_g1=0.96, lambda_ct_islands=0.17, lambda_getis_ord=0.71,
                  lambda_r=2.95e-9, lambda_l2=1e-18)
L2_BYTES = 126e6
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
TRAFFIC = {("c3", "bf16", "gemm_bwd_adam"): 31.11e9,      # profiles/r01_pair_c3_gemm_kernels_raw.csv
           ("c3", "bf16", "gemm_fwd"): 4.22e9, ("c3", "bf16", "gemm_rowdot"): 4.10e9}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def gen_inputs(name, r0, r1, seed=0):
    """Rows [r0, r1) of the synthetic workload (S rows are generated per block so that ranks agree)."""
    N, V, K, T, clusters, _ = WORKLOADS[name]
    if name == "c1":      # real data: the reference's own test fixtures, exported by tests/golden/make_c1_golden.py
        import scipy.sparse as sp
        z = np.load(os.path.join(ROOT, "tests", "golden", "c1_reference.npz"))
        S = sp.csr_matrix((z["S_data"], z["S_indices"], z["S_indptr"]), shape=tuple(z["S_shape"]))[r0:r1].toarray()
        return dict(S=np.ascontiguousarray(S, dtype=np.float32), G=z["G"], d=z["d"])
    rng = np.random.default_rng(seed)
    G = np.log1p(rng.poisson(2.0, (V, K))).astype(np.float32)
    G[:, ~G.any(axis=0)] = 1.0
    out = dict(G=G)
    if clusters:
        w = np.random.default_rng(seed + 1).random(N) + 0.1
        out["d_source"] = (w / w.sum()).astype(np.float32)[r0:r1]
        out["d"] = (np.ones(V) / V).astype(np.float32)
    else:
        out["d"] = (G.sum(axis=1) / G.sum()).astype(np.float32)
    S = np.empty((r1 - r0, K), dtype=np.float32)
    blk = 4096
    for b0 in range((r0 // blk) * blk, r1, blk):
        rb = np.random.default_rng([seed, 7, b0])
        rows = np.log1p(rb.poisson(0.6, (blk, K))).astype(np.float32)
        lo, hi = max(b0, r0), min(b0 + blk, r1)
        S[lo - r0:hi - r0] = rows[lo - b0:hi - b0]
    S[0, ~S.any(axis=0)] = 1.0
    out["S"] = S
    if T:
        lab = np.random.default_rng(seed + 2).integers(0, T, N)[r0:r1]
        E = np.zeros((r1 - r0, T), dtype=np.float32)
        E[np.arange(r1 - r0), lab] = 1.0
        out["ct_encode"] = E
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(name, steps, warmup, target_seconds=20.0):
    """The oracle port of the reference Mapper, timed on the host cores on a bounded sample:
    a row-slice of the cells axis (the work is linear in cells), scaled back to the full workload."""
    import torch
    from oracle.tangram_oracle import OracleMapper
    N, V, K, T, clusters, _ = WORKLOADS[name]
    # ~6*N*V*K flop + ~60 N*V-sized passes per iteration; pick rows for a few seconds per step
    n_s = int(min(N, max(64, 3.0e9 // max(1, V * max(K // 4, 64)))))
    inp = gen_inputs(name, 0, n_s)
    kw = dict(S=inp["S"], G=inp["G"], d=inp["d"], lambda_d=1.0)
    if clusters:
        kw["d_source"] = inp["d_source"] / inp["d_source"].sum()
    M0 = np.random.default_rng(0).standard_normal((n_s, V)).astype(np.float32)
    o = OracleMapper(M0=M0, **kw)
    for _ in range(warmup):
        _, g = o.loss_and_grad(); o.adam_step(g, 0.1)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        _, g = o.loss_and_grad(); o.adam_step(g, 0.1)
        done += 1
        if time.perf_counter() - t0 > target_seconds and done >= 2:
            break
    dt = (time.perf_counter() - t0) / done
    full = dt * (N / n_s)
    return dict(value=1.0 / full, unit="iterations/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{done} steps of the oracle port (torch CPU, closed-form fwd+bwd+Adam) on cells[0:{n_s}] x {V} voxels x {K} genes; "
                       f"{dt * 1e3:.1f} ms/step on the sample, scaled x{N / n_s:.1f} to {N} cells (work is linear in cells); "
                       f"os.cpu_count()={os.cpu_count()}"), dt * (N / n_s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)     # long enough to sit at the sustained (power-capped) clocks
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("TGB200_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("TGB200_PRECISION", "bf16"), choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N, V, K, T, clusters, desc = WORKLOADS[a.workload]
    metric = "map_cells_to_space iterations/sec"
    config = {"workload": f"{desc}, lambda_g1=1, lambda_d=1, lr=0.1", "cells": N, "voxels": V, "genes": K,
              "parallelism": f"cells-sharded x{world}" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        cb, full_dt = cpu_port_rate(a.workload, a.steps, a.warmup, target_seconds=60.0)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "iterations/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": full_dt * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    import torch.distributed as dist
    from tangram_b200 import _lib
    from tangram_b200.engine import Engine
    from tangram_b200.mapping_optimizer import Mapper, shard_rows
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if clusters and world > 1:
        raise SystemExit("clusters mode (c4) does not shard over cells: replicas only (DESIGN.md)")
    r0, r1 = shard_rows(N, rank, world)
    inp = gen_inputs(a.workload, r0, r1)
    lambdas = dict(C5_LAMBDAS) if a.workload == "c5" else {}
    eng = Engine(r1 - r0, V, K, n_types=T, n_cells_global=N, device=local, precision=a.precision,
                 density_mode=_lib.DENSITY_SOURCE if clusters else _lib.DENSITY_CELLS, **lambdas)
    eng.set_expression(inp["S"], inp["G"])
    eng.set_density(inp["d"], inp.get("d_source"))
    graphs = None
    if a.workload == "c5":
        from oracle.tangram_oracle import grid_graph, spatial_weights_from_graph  # input generator only
        conn, dmat = grid_graph(V)
        graphs = {_lib.GRAPH_VOXEL_WEIGHTS: spatial_weights_from_graph(conn, dmat, True, True),
                  _lib.GRAPH_NEIGHBORHOOD_FILTER: spatial_weights_from_graph(conn, dmat, False, False),
                  _lib.GRAPH_SPATIAL_WEIGHTS: spatial_weights_from_graph(conn, dmat, False, True)}
        for which, g in graphs.items():
            eng.set_graph(which, g)
        eng.set_ct_encode(inp["ct_encode"])
    eng.init_mapping_normal(1234 + rank)
    stream = torch.cuda.current_stream().cuda_stream
    xbuf = eng.exchange_tensor() if world > 1 else None

    def one_step():
        if world == 1:
            eng.run(1, 0.1, stream)
        else:
            eng.step_begin(stream)
            dist.all_reduce(xbuf)
            eng.step_end(0.1, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        one_step()
    barrier()
    state_bytes = 3.0 * 4 * N * V / world
    flush = state_bytes < 2 * L2_BYTES
    flush_buf = torch.empty(int(3 * L2_BYTES) // 4, dtype=torch.float32, device="cuda") if flush else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not flush:
        barrier()
        e0.record()
        for _ in range(a.steps):
            one_step()
        e1.record()
        barrier()
        elapsed = e0.elapsed_time(e1) / 1e3
    else:   # small state: L2 flush between timed iterations, per-iteration events
        elapsed = 0.0
        for _ in range(a.steps):
            flush_buf.zero_()
            barrier()
            e0.record()
            one_step()
            e1.record()
            barrier()
            elapsed += e0.elapsed_time(e1) / 1e3
    launches = eng.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    value = a.steps / elapsed

    # ---------------- roofline of the dominant kernel, timed live with CUDA events on this stream
    prof = {}
    for _ in range(3):
        if world == 1:
            for name, ms in eng.profile_step(0.1, stream):
                prof.setdefault(name, []).append(ms)
        else:
            break
    roof = None
    pk = peaks()
    if prof:
        avg = {k: float(np.mean(v)) for k, v in prof.items()}
        step_ms = sum(avg.values())
        top = max(avg, key=avg.get)
        Nl = r1 - r0
        sS = 2.0 if a.precision == "bf16" else (6.0 if a.precision == "bf16x3" else 4.0)
        # algorithmic work of each kernel: flops, HBM bytes per launch
        pb = {"bf16": 2.0, "bf16x3": 6.0}.get(a.precision, 4.0)          # bytes per element of the stored P
        work = {
            "gemm_fwd": (2.0 * Nl * V * K, pb * Nl * V + sS * Nl * K + 4.0 * V * K),
            "gemm_rowdot": (2.0 * Nl * V * K, pb * Nl * V + sS * Nl * K + sS * V * K),
            "gemm_bwd_adam": (2.0 * Nl * V * K, 24.0 * Nl * V + sS * Nl * K + sS * V * K),
            "softmax_rows": (0.0, (4.0 + pb) * Nl * V),
            "loss_reduce": (0.0, 4.0 * V * K * (1 + 1)),
            "scale_rows": (0.0, 6.0 * Nl * K),
        }
        known = {k: v for k, v in avg.items() if any(w in k for w in work)}
        if known:
            top = max(known, key=known.get)
        key = next((k for k in work if k in top), None)
        if key:
            fl, by = work[key]
            t_s = avg[top] / 1e3
            # bf16x3 issues 6 bf16 MMAs per useful product; fp32 FFMA peak: 148 SM x 128 lanes x 2 x 1.965 GHz
            tf_peak = pk["tf_sust"] if a.precision == "bf16" else (pk["tf_sust"] / 6.0 if a.precision == "bf16x3" else 74.0)
            t_fl = fl / (tf_peak * 1e12) if fl else 0.0
            t_by = by / (pk["hbm"] * 1e9)
            if t_fl >= t_by:
                roof = {"bound": "tensor" if a.precision != "fp32" else "fp32-ffma", "achieved": fl / t_s / 1e12,
                        "peak": tf_peak, "unit": "TFLOP/s"}
            else:
                roof = {"bound": "hbm", "achieved": by / t_s / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
            roof["frac"] = roof["achieved"] / roof["peak"]
            roof.update({"kernel": top, "kernel_ms": avg[top], "share_of_step": avg[top] / step_ms, "traffic": TRAFFIC.get((a.workload, a.precision, key)),
                         "peak_source": pk["src"], "per_kernel_ms": avg})
            # the other contraction kernels against the tensor roofline (explains the step)
            others = {}
            for kname, ms in avg.items():
                kk = next((k for k in ("gemm_fwd", "gemm_rowdot", "gemm_bwd_adam") if k in kname), None)
                if kk and a.precision != "fp32":
                    others[kname] = {"tflops": work[kk][0] / (ms / 1e3) / 1e12,
                                     "frac_of_sustained_bf16_peak": work[kk][0] / (ms / 1e3) / 1e12 / pk["tf_sust"],
                                     "hbm_GBs": work[kk][1] / (ms / 1e3) / 1e9}
            roof["contractions"] = others
        hb, fl_it = eng.algorithmic_cost()
        roof_step = max(hb / (pk["hbm"] * 1e9), fl_it / ({"bf16": pk["tf_sust"], "bf16x3": pk["tf_sust"] / 6.0}.get(a.precision, 74.0) * 1e12))
        if roof is not None:
            roof["step_roofline_frac"] = roof_step / (elapsed / a.steps)
            roof["mts_tflops"] = 2.0 * Nl * V * K / (avg.get(next((k for k in avg if "gemm_fwd" in k), top), 1e9) / 1e3) / 1e12

    # ---------------- e2e: the public Mapper API with HOST buffers (H2D + steps + D2H inside the timed region)
    e2e = None
    if not a.no_e2e:
        del eng
        torch.cuda.empty_cache()
        Nl = r1 - r0
        M0 = torch.empty((Nl, V), dtype=torch.float32).pin_memory()
        M0.normal_(generator=torch.Generator().manual_seed(99 + rank))
        Sp = torch.from_numpy(inp["S"]).pin_memory()
        kw = dict(S=Sp.numpy(), G=inp["G"], d=inp["d"], lambda_d=1.0, M0=M0.numpy(), precision=a.precision,
                  device=f"cuda:{local}", n_cells_global=N, process_group=(dist.group.WORLD if world > 1 else None))
        if clusters:
            kw["d_source"] = inp["d_source"]
        if graphs:
            kw.update(lambdas, voxel_weights=graphs[0], neighborhood_filter=graphs[1], spatial_weights=graphs[2],
                      ct_encode=inp["ct_encode"])
        barrier()
        t0 = time.perf_counter()
        mp = Mapper(**kw)
        out, hist = mp.train(a.steps, learning_rate=0.1, print_each=None)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = (Sp.numel() + inp["G"].size + inp["d"].size + M0.numel()) * 4.0
        d2h = out.size * 4.0 + a.steps * 16 * 4.0
        e2e = {"value": a.steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d / a.steps,
               "d2h_bytes_per_step": d2h / a.steps,
               "what": f"Mapper(S,G,d,M0 host).train({a.steps}): upload + {a.steps} iterations + softmax(M) download; "
                       f"total {dt:.2f} s per rank (copies are per call, not per iteration)"}
        del mp

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, _ = cpu_port_rate(a.workload, 6, 1)

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": elapsed / a.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": {"bf16": "bf16 operands / f32 accumulate+state", "bf16x3": "f32 via 3xbf16 split operands on tensor cores / f32 accumulate+state"}.get(a.precision, "f32"),
                "data": "synthetic", "config": dict(config, precision=a.precision,
                                                    l2="L2 flushed between timed iterations" if flush else "state exceeds L2 (no flush)"),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
