#!/usr/bin/env python
"""Benchmark of the RecNet hot path on B200: train samples/s for decoder + local reconstructor (fwd + bwd + clip +
Adam), MSVD shape, batch 100 per GPU, bf16 tensor-core GEMMs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--recon local|global|none]

One process per GPU (torchrun for N > 1).  Prints ONE JSON line on rank 0 (contract in the task statement):
  value      device-timed throughput with inputs resident in HBM (CUDA-graph replay of the whole step)
  e2e        same step driven through the public API from PINNED HOST buffers: H2D of the batch + D2H of the loss
             inside the timed region
  roofline   dominant kernel (by share of the step) timed live with CUDA events around each launch; GEMM-class kernels (the
             tcgen05 GEMMs and the weight-resident persistent loops) are scored against the tensor pipe (burst bf16 peak),
             with the byte-side fraction alongside; `roofline.step` = SURVEY 8(d)'s 397 GFLOP / step time vs the sustained peak
  fp32       the same step in the fp32 parity build (the reference's arithmetic type), value + e2e
  workloads  the other BASELINE configs: global reconstructor, decoder only, greedy decoding at batch 1024, MSR-VTT-shaped stress
  cpu_baseline  the oracle (CPU port of the reference algorithm) timed on this box's host cores, N = 1 only
`--impl reference` times only the CPU oracle (the reference is pure Python and cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = dict(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, cap=30, R=1536)
METRIC = "train samples/s (decoder+local rec fwd+bwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--recon", default="local", choices=["local", "global", "none"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=8)
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32 record and the other-config workloads (N = 1 only anyway)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle (port of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_samples_per_s(recon, iters, warmup=1):
    from oracle import recnet_oracle as O
    s = SHAPE
    torch.set_num_threads(os.cpu_count() or 1)
    feats, targets, masks = O.synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234)
    P = {k: v.requires_grad_(True) for k, v in O.init_decoder_params(s["V"], s["EMB"], s["E"], s["H"], s["A"], seed=0).items()}
    params = list(P.values())
    Q = None
    if recon != "none":
        Q = {k: v.requires_grad_(True) for k, v in O.init_reconstructor_params(recon, s["H"], s["R"], s["A"], seed=1).items()}
    opt_d = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5, amsgrad=True)
    opt_r = torch.optim.Adam(list(Q.values()), lr=1e-6, weight_decay=1e-5) if Q else None

    def step():
        opt_d.zero_grad()
        if opt_r:
            opt_r.zero_grad()
        dl, hid, _, _ = O.forward_decoder(P, feats, targets, masks)
        loss = dl
        if recon == "local":
            loss = loss + O.forward_local_reconstructor(Q, hid, feats)[0]
        elif recon == "global":
            loss = loss + O.forward_global_reconstructor(Q, hid, feats)[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 50.0)
        opt_d.step()
        if opt_r:
            opt_r.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return s["B"] * iters / dt, dt / iters


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    iters = max(1, min(args.steps, 12))
    sps, sec = cpu_oracle_samples_per_s(args.recon, iters, warmup=min(args.warmup, 1) or 1)
    cores = os.cpu_count() or 1
    sample = f"{iters} timed iterations of fwd+bwd+clip+Adam at batch {SHAPE['B']}, L=31, fp32, {cores} torch threads ({cpu_model_name()})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(sps, 3), "unit": "samples/s", "n_gpus": args.gpus, "steps": iters,
        "warmup": 1, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.recon, SHAPE["B"]), "global_batch": SHAPE["B"], "parallelism": "cpu",
                   "implementation": "oracle/recnet_oracle.py (CPU restatement of the reference, pinned to reference-generated fixtures); "
                                     "dropout is evaluated in eval mode (p = 0) by the oracle"},
        "cpu_baseline": {"value": round(sps, 3), "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(sps, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


KCLASS = {1: "gemm_tcgen05", 2: "sgemm_fp32", 3: "attn_fwd", 4: "attn_bwd", 5: "lstm_cell_fwd", 6: "lstm_cell_bwd", 7: "ce_loss",
          8: "splitk_reduce", 9: "persistent_loop", 10: "proj_attn_cell_fwd", 11: "proj_attn_cell_bwd"}


def algorithmic_work(cls, M, N, K, elt):
    """(bound, flops-or-bytes per launch) -- DESIGN.md section 'Kernels and rooflines' states the same formulas."""
    s = SHAPE
    if cls == 9:
        # weight-resident persistent time loop of the local reconstructor (csrc/seq_recon_persist.cuh): (steps, CTAs, 0 fwd / 1 bwd).
        # Per step: the gate GEMM [B x 4R x (H+R)] (fwd) or its transpose dX = dG.W (bwd) + the attention-query GEMM [B x A x R];
        # useful rows only (B of the 112-column UMMA tile)
        B, H, R, A = s["B"], s["H"], s["R"], s["A"]
        if K == 2:       # the decoder's forward loop in 16-CTA clusters (csrc/seq_decoder_cluster.cuh): h_{t-1} . [W_a ; W_hh]^T per step
            return "tensor", M * 2.0 * B * (A + 4 * H) * H
        return "tensor", M * (2.0 * B * 4 * R * (H + R) + 2.0 * B * A * R)
    if cls in (1, 2):
        # useful rows only (M = 100 of the 128-row UMMA tile).  Which roofline binds is decided by the caller from the arithmetic
        # intensity (gemm_bytes): the M = 100 per-step GEMMs move ~90 flop per byte, far below the machine balance (~215).
        return "tensor", 2.0 * M * N * K
    if cls == 3:     # (B, Tn, D): read V + Uv + Wh partials, write ctx
        return "hbm", M * N * K * elt + M * N * s["A"] * 4 + M * s["A"] * 4 + M * K * elt
    if cls == 4:     # read V, dctx partials (~1), Uv ; RMW dUv ; write dWh
        return "hbm", M * N * K * elt + M * K * 4 + 3 * M * N * s["A"] * 4 + 2 * M * s["A"] * 4
    if cls == 5:     # (B, H, n_p): read partials + Gx + c, write gates + c + h (fp32 + operand)
        return "hbm", M * N * (4 * 4 * max(K, 1) + 4 * 4 + 4 + 4 * elt + 4 + 4 + elt)
    if cls == 6:     # read dh terms, gates, c, c_prev, dc; write dG, dc
        return "hbm", M * N * (4 + 4 * max(K, 0) + 4 * elt + 4 + 4 + 4 + 4 * elt + 4)
    if cls == 7:     # (rows, V, fwd/bwd): read logits (+ write dlogits)
        return "hbm", M * N * (4 + (elt if K == 1 else 0))
    if cls == 8:
        return "hbm", M * N * 4 * (K + 1)
    if cls == 10:    # (B, Tn, H) fused projected-feature attention + LSTM cell: read VW, Uv, 4 split-K partial sets of [Wh | gates_h], Gx, c;
        #              write gates stash, c, h (fp32 + operand), e, Wh
        A = s["A"]
        return "hbm", (M * N * 4 * K * elt + M * N * A * 4 + 4 * M * (A + 4 * K) * 4 + M * 4 * K * 4 + M * K * 4
                       + M * 4 * K * elt + M * K * (4 + 4 + elt) + M * (N + A) * 4)
    if cls == 11:    # fused cell backward + attention backward: read VW, Uv, dUv (RMW), 12 dh partial sets, gates, c, c_prev, dc, 2 dh terms;
        #              write [dWh | dG] operand row, dc, dWh, dw
        A = s["A"]
        return "hbm", (M * N * 4 * K * elt + 3 * M * N * A * 4 + 12 * M * K * 4 + M * 4 * K * elt + 5 * M * K * 4
                       + M * (A + 4 * K) * elt + M * K * 4 + 3 * M * A * 4)
    return "hbm", 0.0


def gemm_bytes(M, N, K, elt):
    """Algorithmic bytes of one GEMM launch: both operands once + the fp32 result once."""
    return (M * K + N * K) * elt + M * N * 4.0


# ncu --set full captures committed under profiles/ (tools/collect_profiles.sh + tools/summarise_profiles.py): DRAM traffic per
# launch of the hot-loop kernels, keyed by the (kernel class, shape) the live profile reports
NCU_KEYS = {("gemm_tcgen05", (100, 6144, 2048)): "gate_gemm_warm", ("proj_attn_cell_fwd", (100, 28, 512)): "pf_fwd_kernel",
            ("proj_attn_cell_bwd", (100, 28, 512)): "pf_bwd_kernel", ("attn_fwd", (100, 31, 512)): "lean_fwd_kernel",
            ("attn_bwd", (100, 31, 512)): "lean_bwd_kernel", ("lstm_cell_fwd", (100, 1536, 3)): "lstm_cell_fwd_kernel",
            ("lstm_cell_bwd", (100, 1536, 11)): "lstm_cell_bwd_kernel",
            # r2: the weight-resident persistent loops (shape = steps, CTAs, 0 fwd / 1 bwd / 2 decoder cluster loop)
            ("persistent_loop", (28, 144, 0)): "local_fwd_kernel", ("persistent_loop", (28, 144, 1)): "local_bwd_kernel",
            ("persistent_loop", (31, 112, 2)): "decoder_fwd_cluster_kernel"}


def ncu_traffic(kernel, shape):
    """(steady-state traffic bytes, cold-cache traffic bytes or None, source file) from the newest committed ncu summary."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_ncu_kernels.json")))
    key = NCU_KEYS.get((kernel, tuple(shape)))
    if not files or key is None:
        return None, None, None
    try:
        d = json.load(open(files[-1]))
        warm = d.get(key, {}).get("traffic_bytes")
        cold = d.get(key.replace("_warm", "_cold"), {}).get("traffic_bytes") if key.endswith("_warm") else None
        return warm, cold, os.path.basename(files[-1])
    except Exception:
        return None, None, None


def workload_text(recon, batch):
    return (f"RecNet decoder + {recon} reconstructor train step (fwd+bwd+clip+Adam), MSVD shape: 28x1536 InceptionV4 feats, "
            f"caption len 30 (L=31 steps), emb 468, attn 128, hidden 512, rec hidden 1536, vocab 4188, batch {batch} per GPU, dropout on")


# ------------------------------------------------------------------------------------------------------------
# the other BASELINE configs and the fp32 build: compact sub-records (N = 1, inputs resident, CUDA-graph replay)
# ------------------------------------------------------------------------------------------------------------
def _configure(C, shp, recon, precision, dec_layers, device):
    C.decoder_model = C.reconstructor_model = "LSTM"
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = shp["B"], shp["cap"], shp["T"], shp["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size, C.embedding_size = dec_layers, shp["H"], shp["A"], shp["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = 1, shp["R"], shp["A"]
    C.use_recon = recon != "none"
    C.reconstructor_type = recon if C.use_recon else "local"
    C.precision, C.device = precision, device


def _graph_time(fn, steps, warmup, grad=True):
    """ms per call of fn under CUDA-graph replay (eager if capture fails -- reported)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.set_grad_enabled(grad):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph, run = None, fn
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.set_grad_enabled(grad):
            fn()
        run = graph.replay
    except Exception as ex:
        print(f"[bench] sub-workload graph capture failed ({type(ex).__name__}: {ex}); timing eager launches", file=sys.stderr)
        graph = None
        torch.cuda.synchronize()
    with torch.set_grad_enabled(grad):
        for _ in range(max(3, warmup)):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, graph is not None


def _top_kernel(lib, fn, shp, elt, pk, grad=True):
    """Dominant kernel of one eager call of fn (same event-pair profile leg as the headline record)."""
    import ctypes
    with torch.set_grad_enabled(grad):
        lib.recnet_profile_enable(1, 4096)
        fn()
        lib.recnet_profile_enable(0, 0)
    buf = (ctypes.c_float * (4096 * 5))()
    nrec = lib.recnet_profile_collect(ctypes.cast(buf, ctypes.c_void_p), 4096)
    agg = {}
    for i in range(max(nrec, 0)):
        key = (int(buf[5 * i]), int(buf[5 * i + 1]), int(buf[5 * i + 2]), int(buf[5 * i + 3]))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1; a[1] += buf[5 * i + 4]
    if not agg:
        return None
    tot = sum(v[1] for v in agg.values())
    (cls, M, N, K), (cnt, ms) = max(agg.items(), key=lambda kv: kv[1][1])
    global SHAPE
    saved, SHAPE = SHAPE, dict(SHAPE, **shp)
    try:
        bound, work = algorithmic_work(cls, M, N, K, elt)
    finally:
        SHAPE = saved
    avg_s = ms / cnt * 1e-3
    peak = pk["tf_burst"] * 1e12 if bound == "tensor" else pk["hbm"] * 1e9
    return {"kernel": KCLASS.get(cls, "other"), "shape": [M, N, K], "launches": cnt, "avg_us": round(ms / cnt * 1e3, 2),
            "share": round(ms / tot, 4), "bound": bound, "frac": round(work / avg_s / peak, 4) if work else None}


def extra_records(args, dev, lib, T, synthetic_batch, pk):
    """fp32 build of the headline step + BASELINE configs 1 (decoder only), 2 (global reconstructor), 4 (greedy, batch 1024) and
    5 (MSR-VTT-shaped stress: 40 frames x 3584-d, 2-layer LSTM decoder, local reconstructor with R = 3584), batch 100 on this GPU."""
    C = T.C
    steps, warm = max(10, min(args.steps, 20)), 3
    out = {}

    def train_record(shp, recon, precision, dec_layers=1, with_e2e=False):
        _configure(C, shp, recon, precision, dec_layers, str(dev))
        torch.manual_seed(0)
        dec = T.build_decoder(shp["V"])
        rec = T.build_reconstructor() if recon != "none" else None
        L = shp["cap"] + 1
        fh, th, _ = synthetic_batch(shp["B"], shp["T"], shp["E"], shp["V"], shp["cap"], seed=1234)
        fh, th = fh.pin_memory(), th.pin_memory()
        f, t = fh.to(dev), th.to(dev)
        loss_d = torch.zeros((), device=dev)

        def step():
            loss, _, _ = T.train_step(dec, rec, f, t, n_steps=L)
            loss_d.copy_(loss.detach())
        ms, graphed = _graph_time(step, steps, warm)
        r = {"value": round(shp["B"] / (ms * 1e-3), 2), "unit": "samples/s", "ms_per_step": round(ms, 4), "dtype": precision,
             "cuda_graph": graphed, "steps": steps}
        n0 = lib.recnet_launch_count()
        step()
        r["launches_per_step"] = int(lib.recnet_launch_count() - n0)
        r["top_kernel"] = _top_kernel(lib, step, shp, 2 if precision == "bf16" else 4, pk)
        if with_e2e:      # host buffers -> H2D -> step -> D2H loss, every step inside the timed region (no overlap tricks)
            lh = torch.zeros(1).pin_memory()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                f.copy_(fh, non_blocking=True); t.copy_(th, non_blocking=True)
                step()
                lh.copy_(loss_d.view(1), non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ems = e0.elapsed_time(e1) / steps
            r["e2e"] = {"value": round(shp["B"] / (ems * 1e-3), 2), "unit": "samples/s", "ms_per_step": round(ems, 4),
                        "h2d_bytes_per_step": fh.numel() * 4 + th.numel() * 8, "d2h_bytes_per_step": 4,
                        "pipeline": "pinned host -> H2D -> eager launches of the step -> D2H loss (no overlap, no graph)"}
        del dec, rec
        torch.cuda.empty_cache()
        return r

    s = SHAPE
    try:
        out["fp32"] = dict(train_record(s, args.recon, "fp32", with_e2e=True),
                           note="same step in the fp32 parity build: fp32 storage, FFMA GEMMs, libm activations (1e-3 vs the oracle)")
    except Exception as ex:
        out["fp32"] = {"error": f"{type(ex).__name__}: {ex}"}
    wl = {}
    for name, (shp, recon, layers, text) in {
        "dec_only": (s, "none", 1, "config 1 shape on the GPU: decoder only (wo. reconstructor), batch 100"),
        "global": (s, "global", 1, "config 2: RecNet + global reconstructor, bf16, batch 100"),
        "msrvtt_stress": (dict(s, T=40, E=3584, R=3584), "local", 2,
                          "config 5 per-GPU share: 40 frames x (1536 + 2048)-d features, 2-layer LSTM decoder, local reconstructor, batch 100"),
    }.items():
        try:
            wl[name] = dict(train_record(shp, recon, "bf16", layers), workload=text)
        except Exception as ex:
            wl[name] = {"error": f"{type(ex).__name__}: {ex}", "workload": text}
    # config 4: greedy decoding, decoder only, batch 1024, 28 frames, max len 30 (eval.greedy_search -> one C call, argmax feedback on device)
    try:
        _configure(C, dict(s, B=1024), "none", "bf16", 1, str(dev))
        torch.manual_seed(0)
        dec = T.build_decoder(s["V"])
        dec["model"].eval()
        f, _, _ = synthetic_batch(1024, s["T"], s["E"], s["V"], s["cap"], seed=77)
        f = f.to(dev)
        L = s["cap"] + 1
        fn = lambda: dec["model"].greedy(f, L)
        ms, graphed = _graph_time(fn, steps, warm, grad=False)
        n0 = lib.recnet_launch_count()
        fn()
        wl["greedy_b1024"] = {"value": round(1024 / (ms * 1e-3), 2), "unit": "captions/s", "ms_per_batch": round(ms, 4), "dtype": "bf16",
                              "cuda_graph": graphed, "steps": steps, "launches_per_batch": int(lib.recnet_launch_count() - n0),
                              "gflop": 481.7, "tflops": round(481.7e9 / (ms * 1e-3) / 1e12, 2),
                              "top_kernel": _top_kernel(lib, fn, dict(s, B=1024), 2, pk, grad=False),
                              "workload": "config 4: greedy inference, decoder only, batch 1024, 28 frames, 31 steps (no early stop with random weights)"}
        del dec
        torch.cuda.empty_cache()
    except Exception as ex:
        wl["greedy_b1024"] = {"error": f"{type(ex).__name__}: {ex}"}
    out["workloads"] = wl
    return out


def dbg(rank, msg):
    if os.environ.get("RECNET_BENCH_VERBOSE"):
        print(f"[bench rank {rank} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    # the contract is ONE JSON line on stdout: park fd 1 on stderr while libraries initialise (NCCL prints its version
    # banner to stdout) and restore it just before printing the result
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    import recnet_b200
    from recnet_b200 import _lib as RL, train as T
    from recnet_b200.data import synthetic_batch
    from recnet_b200.parallel import broadcast_parameters, make_reducer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dp_self = world == 1 and os.environ.get("RECNET_DP_SELF") == "1"   # developer probe: 1-rank NCCL group, collective captured in the graph
    if dp_self:
        for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29977"), ("RANK", "0"), ("WORLD_SIZE", "1")):
            os.environ.setdefault(k, v)
    if world > 1 or dp_self:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # whole-step CUDA-graph capture includes the NCCL all-reduces: the process group's watchdog must not poll
        # events while a capture is open (PyTorch's documented requirement for DDP + graph capture)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev)
    RL.require_device(local_rank)
    s = SHAPE
    C = T.C
    C.decoder_model = C.reconstructor_model = "LSTM"
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = s["B"], s["cap"], s["T"], s["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size, C.embedding_size = 1, s["H"], s["A"], s["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = 1, s["R"], s["A"]
    C.use_recon = args.recon != "none"
    C.reconstructor_type = args.recon if C.use_recon else "local"
    C.precision, C.device = args.precision, f"cuda:{local_rank}"
    torch.manual_seed(0)
    dec = T.build_decoder(s["V"])
    rec = T.build_reconstructor() if C.use_recon else None
    modules = [dec["model"]] + ([rec["model"]] if rec else [])
    broadcast_parameters(modules)
    # reconstructor first: its gradients are final first (its backward runs before the decoder's BPTT)
    reducer = make_reducer(([rec["model"]] if rec else []) + [dec["model"]])
    L = s["cap"] + 1

    # synthetic MSVD-shaped batch of this rank's shard, in pinned host memory (the e2e leg copies from here every step)
    feats_h, targets_h, _ = synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234 + rank)
    feats_h, targets_h = feats_h.pin_memory(), targets_h.pin_memory()
    feats_d = torch.empty_like(feats_h, device=dev)
    targets_d = torch.empty_like(targets_h, device=dev)
    feats_d.copy_(feats_h); targets_d.copy_(targets_h)
    loss_d = torch.zeros((), device=dev)

    def make_step(fd, td):
        def _step():
            reducer.start_iteration()
            loss, _, _ = T.train_step(dec, rec, fd, td, n_steps=L, reducer=reducer if world > 1 or dp_self else None)
            loss_d.copy_(loss.detach())
        return _step

    step = make_step(feats_d, targets_d)

    lib = RL.lib()
    dbg(rank, "models built, starting eager warm-up")
    # ---- warm-up (eager), launch count, per-kernel timing leg -------------------------------------------------
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
        side.synchronize()
        n0 = lib.recnet_launch_count()
        step()
        launches_per_step = lib.recnet_launch_count() - n0
        side.synchronize()
        # per-launch CUDA-event timing of one eager step (same kernels, same shapes, same stream as the timed region)
        lib.recnet_profile_enable(1, 4096)
        step()
        lib.recnet_profile_enable(0, 0)
        import ctypes
        buf = (ctypes.c_float * (4096 * 5))()
        nrec = lib.recnet_profile_collect(ctypes.cast(buf, ctypes.c_void_p), 4096)
    torch.cuda.current_stream().wait_stream(side)
    dbg(rank, f"eager warm-up + profile leg done ({launches_per_step} launches/step)")
    elt = 2 if args.precision == "bf16" else 4
    agg = {}
    for i in range(max(nrec, 0)):
        cls, M, N, K, ms = int(buf[5 * i]), int(buf[5 * i + 1]), int(buf[5 * i + 2]), int(buf[5 * i + 3]), buf[5 * i + 4]
        a = agg.setdefault((cls, M, N, K), [0, 0.0])
        a[0] += 1; a[1] += ms
    prof_total_ms = sum(v[1] for v in agg.values())

    # ---- CUDA graph of the whole step -----------------------------------------------------------------------
    graph = None
    if not args.no_graph:
        try:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            dump = os.environ.get("RECNET_GRAPH_DUMP")           # developer probe: DOT file of the captured step graph
            if dump:
                graph.enable_debug_mode()
            with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=int(os.environ.get("RECNET_CAPTURE_PRIO", "0"))),
                                  capture_error_mode="thread_local" if (world > 1 or dp_self) else "global"):
                step()
            if dump and rank == 0:
                graph.debug_dump(dump)
        except Exception as ex:                     # report, never silently change what is measured
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed ({type(ex).__name__}: {ex}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    dbg(rank, f"graph captured: {graph is not None}")
    # e2e leg, single GPU: a second capture of the same step over a second pair of static input buffers (same memory pool), so
    # that step i+1's batch can be copied from pinned host memory STRAIGHT into its graph's inputs while step i runs
    # (no device-to-device hop through a staging buffer: 34 MB less L2 / HBM traffic per step).
    graph_b, feats_b, targets_b = None, None, None
    two_graphs = world == 1 or os.environ.get("RECNET_BENCH_TWO_GRAPHS", "0") == "1"
    if graph is not None and two_graphs:
        try:
            feats_b, targets_b = feats_d.clone(), targets_d.clone()
            step_b = make_step(feats_b, targets_b)
            graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_b, pool=graph.pool(), stream=torch.cuda.Stream(priority=int(os.environ.get("RECNET_CAPTURE_PRIO", "0"))),
                                  capture_error_mode="thread_local" if (world > 1 or dp_self) else "global"):
                step_b()
        except Exception as ex:
            print(f"[bench] second capture for the e2e leg failed ({type(ex).__name__}: {ex}); using the staging-buffer pipeline", file=sys.stderr)
            graph_b = None
            torch.cuda.synchronize()
    if world > 1:      # all ranks must run the same mode, or the collectives would not match up
        flag = torch.tensor([1 if graph is not None else 0, 1 if graph_b is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag[0]) == 0:
            graph = None
        if int(flag[1]) == 0:
            graph_b = None
    # The timed region replays the captured step (one graph exec, back to back).  RECNET_BENCH_ALT=1 alternates between the two
    # captures instead (same step, two static input buffers): measured identical on the B200 (2.836 vs 2.839 ms), so re-launching
    # one exec costs nothing extra and the simpler loop stays the default.
    replays = [g.replay for g in ((graph, graph_b) if os.environ.get("RECNET_BENCH_ALT", "0") == "1" else (graph,)) if g is not None]
    if graph is None:
        run = lambda i=0: step()
    else:
        run = lambda i=0: replays[i % len(replays)]()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for i in range(args.warmup):
        run(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dbg(rank, "warm-up replays done")
    total_ms = timed(run, args.steps)
    clocks = sampler.stop() if sampler else None
    dbg(rank, f"timed region done: {total_ms:.2f} ms")

    # ---- e2e: host buffers -> H2D -> step -> D2H loss, every step, inside the timed region ------------------------
    # Software pipeline: step i+1's batch is copied from pinned host memory into a staging buffer on a copy stream
    # while step i computes; the step itself starts with a 17 MB device-to-device hop (~6 us) into the graph's
    # static input.  Every step's H2D copy and D2H loss read are issued inside the timed region.
    losses_h = torch.zeros(args.steps + 2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    stage_f = [torch.empty_like(feats_d) for _ in range(2)]
    stage_t = [torch.empty_like(targets_d) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])
            stage_f[k].copy_(feats_h, non_blocking=True)
            stage_t[k].copy_(targets_h, non_blocking=True)
            ev_ready[k].record(copy_stream)

    in_f, in_t, runs = [feats_d, feats_b], [targets_d, targets_b], [graph.replay if graph is not None else step,
                                                                     graph_b.replay if graph_b is not None else None]

    def issue_h2d_direct(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])            # the step that last read buffer k has finished
            in_f[k].copy_(feats_h, non_blocking=True)
            in_t[k].copy_(targets_h, non_blocking=True)
            ev_ready[k].record(copy_stream)

    def e2e_step_direct(i):
        k = i & 1
        if i == 0:
            issue_h2d_direct(0)
        main = torch.cuda.current_stream()
        issue_h2d_direct(k ^ 1)              # next step's batch -> the other graph's inputs, overlapped with this step
        main.wait_event(ev_ready[k])
        runs[k]()
        ev_free[k].record(main)
        losses_h[i:i + 1].copy_(loss_d.view(1), non_blocking=True)

    def e2e_step_staged(i):
        k = i & 1
        if i == 0:
            issue_h2d(0)
        main = torch.cuda.current_stream()
        main.wait_event(ev_ready[k])
        feats_d.copy_(stage_f[k], non_blocking=True)
        targets_d.copy_(stage_t[k], non_blocking=True)
        ev_free[k].record(main)
        issue_h2d(k ^ 1)                     # next step's batch, overlapped with this step's compute
        runs[0]()
        losses_h[i:i + 1].copy_(loss_d.view(1), non_blocking=True)

    e2e_step = e2e_step_direct if graph_b is not None else e2e_step_staged
    for k in range(2):
        ev_free[k].record(torch.cuda.current_stream())
    dbg(rank, "clock sampler stopped; e2e warm-up")
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    for k in range(2):
        ev_free[k].record(torch.cuda.current_stream())
    dbg(rank, "e2e warm-up issued")
    e2e_ms = timed(e2e_step, args.steps)
    dbg(rank, f"e2e region done: {e2e_ms:.2f} ms")
    h2d = feats_h.numel() * 4 + targets_h.numel() * 8
    d2h = 4

    samples = s["B"] * world * args.steps
    value = samples / (total_ms * 1e-3)
    e2e_value = samples / (e2e_ms * 1e-3)

    if rank == 0:
        pk = peaks()
        kernels = []
        for (cls, M, N, K), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            bound, work = algorithmic_work(cls, M, N, K, elt)
            avg_s = ms / cnt * 1e-3
            extra = {}
            if cls in (1, 2, 9):
                # GEMM-class kernels are scored on the tensor pipe against the BURST bf16 peak (each is a 5-500 us
                # kernel, not a seconds-long loop); the byte-side fraction (operands once + result once over the copy bandwidth)
                # is kept alongside: the M = 100 per-step shapes move ~90 flop per byte, below the machine balance
                nbytes = gemm_bytes(M, N, K, elt) if cls != 9 else M * gemm_bytes(s["B"], 4 * s["R"], s["H"] + s["R"], elt)
                extra = {"flop_per_byte": round(work / nbytes, 1), "tensor_frac": round(work / avg_s / 1e12 / pk["tf_burst"], 4),
                         "hbm_frac": round(nbytes / avg_s / 1e9 / pk["hbm"], 4)}
            if bound == "tensor":
                ach, peak, unit = work / avg_s / 1e12, pk["tf_burst"], "TFLOP/s"
            else:
                ach, peak, unit = work / avg_s / 1e9, pk["hbm"], "GB/s"
            kernels.append({"kernel": KCLASS.get(cls, "other"), "shape": [M, N, K], "launches": cnt, "avg_us": round(ms / cnt * 1e3, 2),
                            "share": round(ms / prof_total_ms, 4) if prof_total_ms else None, "bound": bound,
                            "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4), **extra})
        top = kernels[0] if kernels else None
        roofline = None
        if top:
            tr_warm, tr_cold, tr_src = ncu_traffic(top["kernel"], top["shape"])
            roofline = {"bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                        "traffic": tr_warm, "traffic_cold_cache": tr_cold, "traffic_source": tr_src,
                        "kernel": top["kernel"], "shape_MNK": top["shape"], "avg_us": top["avg_us"],
                        "share_of_profiled_step": top["share"], "peak_source": pk["src"] + (" (burst bf16)" if top["bound"] == "tensor" else " (copy bandwidth)"),
                        "note": "avg_us is a per-launch CUDA-event pair in an eager step (adds ~4 us to a launch: < 1 % of a persistent loop, "
                                "up to 40 % of a 10 us per-step kernel -- compare with the committed ncu launch list under profiles/); traffic = dram "
                                "bytes of one ncu --set full launch in steady state; GEMM-class kernels: bound = tensor pipe (SURVEY 8d), frac vs the "
                                "burst bf16 peak, hbm_frac = operands-once + result-once bytes vs the copy bandwidth",
                        "step": {"gflop": 397.0, "tflops": round(397.0e9 / (total_ms / args.steps * 1e-3) / 1e12, 2),
                                 "frac": round(397.0e9 / (total_ms / args.steps * 1e-3) / 1e12 / pk["tf_sust"], 4),
                                 "peak": pk["tf_sust"], "peak_source": pk["src"] + " (sustained bf16)",
                                 "note": "SURVEY 8(d): 397.0 GFLOP per iteration of decoder + local reconstructor, fwd + bwd, batch 100 (config 3)"}
                        if args.recon == "local" else None}
            for k in ("flop_per_byte", "tensor_frac", "hbm_frac"):
                if k in top:
                    roofline[k] = top[k]
        cpu = None
        if world == 1 and args.cpu_iters > 0:
            sps, sec = cpu_oracle_samples_per_s(args.recon, args.cpu_iters)
            cores = os.cpu_count() or 1
            cpu = {"value": round(sps, 3), "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_iters} iterations of fwd+bwd+clip+Adam, batch {s['B']}, L=31, fp32 oracle, {cores} torch threads ({cpu_model_name()}), {sec * 1e3:.0f} ms/iter"}
        extras = {}
        if world == 1 and not args.no_extras and not args.no_graph:
            dbg(rank, "extra records (fp32 build, other configs)")
            extras = extra_records(args, dev, lib, T, synthetic_batch, pk)
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_text(args.recon, s["B"]),
                       "global_batch": s["B"] * world, "parallelism": f"dp{world}", "cuda_graph": graph is not None,
                       "graph_execs": len(replays) if graph is not None else 0,
                       "l2": "no explicit flush: every step streams ~0.9 GB of weights/optimizer state/activation stash, >> 126 MB L2"},
            "e2e": {"value": round(e2e_value, 2), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms / args.steps, 4),
                    "pipeline": ("pinned host -> H2D straight into the inputs of one of two captured step graphs (alternating), overlapped with the previous step"
                                 if graph_b is not None else "pinned host -> H2D into a staging buffer (overlapped), device copy into the graph inputs")},
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline, "kernels": kernels[:12], "cpu_baseline": cpu,
            "allreduce_bytes_per_step": reducer.bytes_last, "allreduce_impl": type(reducer).__name__,
            **extras,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        dist.barrier()
        torch.cuda.synchronize()
        if type(reducer).__name__ == "NvlinkAllReducer":
            # no NCCL work sits in the captured graphs (the gradient exchange is our own kernel): a regular teardown works
            del graph, graph_b
            reducer.remove()
            dist.destroy_process_group()
        else:
            # RECNET_DP_IMPL=nccl: with NCCL work captured inside a live CUDA graph destroy_process_group() blocks (observed on the
            # 2-GPU box); everything is flushed, so exit directly
            os._exit(0)


if __name__ == "__main__":
    main()
