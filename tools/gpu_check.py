"""Developer diagnostics for the GPU box: run every parity check, print the error of every quantity, keep going
after a failure.  (The pass/fail gates are in tests/ -m gpu; this prints numbers for debugging.)"""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import recnet_b200                                   # noqa: E402
from recnet_b200 import _lib as L, ops               # noqa: E402
from recnet_b200 import functional as Fn             # noqa: E402
from recnet_b200 import train as T                   # noqa: E402
from oracle import recnet_oracle as O                # noqa: E402
from tests.golden_util import load_golden            # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def section(name):
    print(f"\n=== {name} ===", flush=True)


def run(fn, *a, **k):
    try:
        t0 = time.time()
        fn(*a, **k)
        torch.cuda.synchronize()
        print(f"  [{fn.__name__} ok, {time.time() - t0:.2f}s]", flush=True)
    except Exception:
        traceback.print_exc()
        print(f"  [{fn.__name__} FAILED]", flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("  CUDA context is broken:", e)
            sys.exit(3)


def check_gemm(prec):
    dt = torch.bfloat16 if prec == L.PREC_BF16 else torch.float32
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [(100, 2048, 2048), (128, 128, 64), (100, 128, 512), (3100, 4188, 512), (300, 72, 200), (2048, 468, 3100),
              (128, 1536, 2800)]
    for (M, N, K) in shapes:
        for tA in (False, True):
            for tB in (False, True):
                if prec == L.PREC_BF16 and (((M if tA else K) % 8) or ((N if tB else K) % 8)):
                    continue
                A = torch.randn((K, M) if tA else (M, K), generator=g).to(dev).to(dt)
                B = torch.randn((K, N) if tB else (N, K), generator=g).to(dev).to(dt)
                bias = torch.randn(N, generator=g).to(dev)
                ref = (A.double().t() if tA else A.double()) @ (B.double() if tB else B.double().t()) + bias.double()
                for splits in (1, 3):
                    for bn in ((0,) if prec == L.PREC_FP32 else (64, 128)):
                        try:
                            out = ops.gemm(prec, A, tA, B, tB, bias=bias, splits=splits, bn_hint=bn)
                            if splits > 1:
                                out = out.sum(0)
                            torch.cuda.synchronize()
                            e = rel(out, ref)
                            flag = "" if e < (2e-2 if prec else 1e-4) else "   <<<<<< BAD"
                            print(f"  gemm prec={prec} M{M} N{N} K{K} tA={int(tA)} tB={int(tB)} splits={splits} bn={bn}: rel {e:.2e}{flag}", flush=True)
                        except Exception as ex:
                            print(f"  gemm prec={prec} M{M} N{N} K{K} tA={int(tA)} tB={int(tB)} splits={splits} bn={bn}: EXC {ex}", flush=True)
                            torch.cuda.synchronize()


def build_models(g, kind, precision):
    m = g["meta"]
    C = T.C
    C.decoder_model, C.reconstructor_model = m["dec_model"], m["rec_model"]
    C.batch_size, C.caption_max_len = m["B"], m["cap_len"]
    C.encoder_output_len, C.encoder_output_size = m["T"], m["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size = m["dec_layers"], m["H"], m["A"]
    C.embedding_size = m["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = m["rec_layers"], m["E"], m["A"]
    C.precision = precision
    C.device = "cuda"
    dec = T.build_decoder(m["V"])
    dec["model"].load_state_dict({k: v.float() for k, v in g["dec"].items()})
    dec["model"].eval()
    rec = None
    if kind != "none":
        C.reconstructor_type = kind
        rec = T.build_reconstructor()
        rec["model"].load_state_dict({k: v.float() for k, v in g[kind].items()})
        rec["model"].eval()
    return dec, rec


def check_golden(name, precision):
    g = load_golden(name)
    feats = g["feats"].float().to(dev)
    targets = g["targets"].to(dev)
    masks = targets > 0
    for kind in ("none", "global", "local"):
        dec, rec = build_models(g, kind, precision)
        dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, masks, 1.0)
        loss = dloss
        line = f"  {name} {precision} {kind}: dec_loss rel {rel(dloss, torch.tensor(g['dec_loss'])):.2e} hid {rel(hiddens, g['hiddens']):.2e}"
        if rec is not None:
            fwd = T.forward_global_reconstructor if kind == "global" else T.forward_local_reconstructor
            rloss = fwd(hiddens, feats, rec)
            line += f" rec_loss rel {rel(rloss, torch.tensor(g[kind + '_loss'])):.2e}"
            loss = dloss + 1.0 * rloss
        print(line, flush=True)
        loss.backward()
        try:
            Fn.check_loop_status()
        except RuntimeError as ex:
            print("   LOOP STATUS:", ex, flush=True)
        worst = 0.0
        for k, ref in g["grads"][kind].items():
            owner, key = k.split(".", 1)
            mod = dec["model"] if owner == "dec" else rec["model"]
            p = dict(mod.named_parameters())[key]
            e = rel(p.grad, ref)
            worst = max(worst, e)
            print(f"      grad {k:34s} rel {e:.2e}", flush=True)
        print(f"    worst grad rel {worst:.2e}", flush=True)
    # greedy
    dec, _ = build_models(g, "none", precision)
    ids, n = dec["model"].greedy(feats, g["meta"]["cap_len"] + 1)
    n = int(n.item())
    same = n == g["greedy_ids"].shape[0] and torch.equal(ids[:n].cpu(), g["greedy_ids"])
    print(f"  {name} {precision} greedy: steps {n} vs {g['greedy_ids'].shape[0]} ids equal: {same}", flush=True)
    # per-step module forward
    B, H = feats.shape[0], g["meta"]["H"]
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev)
    hid = (torch.zeros(1, B, H, device=dev), torch.zeros(1, B, H, device=dev))
    with torch.no_grad():
        logits, _ = dec["model"](tok, hid, feats)
    print(f"  {name} {precision} Decoder.forward step0 logits rel {rel(logits, g['step0_logits']):.2e}", flush=True)


def check_full(precision, kind="local", B=100):
    C = T.C
    C.decoder_model = C.reconstructor_model = "LSTM"
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = B, 30, 28, 1536
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size, C.embedding_size = 1, 512, 128, 468
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = 1, 1536, 128
    C.reconstructor_type, C.precision, C.device = kind, precision, "cuda"
    V = 4188
    feats, targets, masks = O.synthetic_batch(B, 28, 1536, V, 30, seed=1234)
    P = O.init_decoder_params(V, 468, 1536, 512, 128, seed=0)
    Q = O.init_reconstructor_params(kind, 512, 1536, 128, seed=1)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    Qr = {k: v.clone().requires_grad_(True) for k, v in Q.items()}
    t0 = time.time()
    dl, hid, _, aux = O.forward_decoder(Pr, feats, targets, masks)
    if kind == "local":
        rl, _ = O.forward_local_reconstructor(Qr, hid, feats)
    else:
        rl, _ = O.forward_global_reconstructor(Qr, hid, feats)
    (dl + rl).backward()
    print(f"  oracle fp32 CPU fwd+bwd {time.time() - t0:.1f}s  dec {dl.item():.6f} rec {rl.item():.6f}", flush=True)
    dec = T.build_decoder(V)
    dec["model"].load_state_dict(P)
    rec = T.build_reconstructor()
    rec["model"].load_state_dict(Q)
    dec["model"].eval(); rec["model"].eval()
    f, t, m = feats.to(dev), targets.to(dev), masks.to(dev)
    dloss, hiddens, _ = T.forward_decoder(dec, f, t, m, 1.0)
    fwd = T.forward_global_reconstructor if kind == "global" else T.forward_local_reconstructor
    rloss = fwd(hiddens, f, rec)
    (dloss + rloss).backward()
    torch.cuda.synchronize()
    try:
        Fn.check_loop_status()
    except RuntimeError as ex:
        print("   LOOP STATUS:", ex, flush=True)
    print(f"  full {precision} {kind}: dec_loss rel {rel(dloss, dl):.2e} rec_loss rel {rel(rloss, rl):.2e} hiddens rel {rel(hiddens[:, 0], hid[:, 0]):.2e}", flush=True)
    for k, p in dec["model"].named_parameters():
        print(f"      grad dec.{k:28s} rel {rel(p.grad, Pr[k].grad):.2e}", flush=True)
    for k, p in rec["model"].named_parameters():
        print(f"      grad rec.{k:28s} rel {rel(p.grad, Qr[k].grad):.2e}", flush=True)
    # timing (eager, no graph)
    for _ in range(2):
        T.train_step(dec, rec, f, t, n_steps=31)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        T.train_step(dec, rec, f, t, n_steps=31)
    e1.record(); torch.cuda.synchronize()
    print(f"  full {precision} {kind}: eager train_step {e0.elapsed_time(e1) / 5:.3f} ms/iter", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm32", "gemm16", "golden32", "golden16", "full32", "full16"]
    print(torch.cuda.get_device_name(0), flush=True)
    L.require_device(0)
    if "gemm32" in what:
        section("sgemm fp32"); run(check_gemm, L.PREC_FP32)
    if "golden32" in what:
        section("golden fp32")
        for n in ("tiny_lstm", "tiny_lstm_ragged", "small_lstm"):
            run(check_golden, n, "fp32")
    if "gemm16" in what:
        section("tcgen05 gemm bf16"); run(check_gemm, L.PREC_BF16)
    if "golden16" in what:
        section("golden bf16")
        for n in ("tiny_lstm", "small_lstm"):
            run(check_golden, n, "bf16")
    if "full32" in what:
        section("full size fp32"); run(check_full, "fp32", "local")
    if "full16" in what:
        section("full size bf16"); run(check_full, "bf16", "local"); run(check_full, "bf16", "global")
