"""On-GPU kernel self-check: every C-ABI op against a torch fp64 reference, one subprocess per group so that a
device trap in one group does not poison the rest.  Prints one line per case; exit code = number of failed groups.

    python tools/gpu_selfcheck.py [group ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_f32", "gemm_bf16_kk", "gemm_bf16_mn", "gemm_bf16_batch", "rowops", "token_mix", "chain_f32", "chain_fwd",
          "chain_bwd", "chain_unfused", "linear", "heads", "adam", "patch_embed", "gemm_bf16_wide"]


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, err, tol):
    ok = err == err and err < tol
    print(f"  [{'ok' if ok else 'FAIL'}] {name}: rel_err={err:.3e} (tol {tol:.0e})", flush=True)
    return ok


def run_group(g):
    import torch
    from m2_mixer_b200 import ops
    from m2_mixer_b200._lib import BF16, FP32
    from oracle import m2mixer_oracle as O
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    rn = lambda *s: torch.randn(*s, device=dev, dtype=torch.float32)

    def gemm_case(prec, M, N, K, a_mn, b_mn, **kw):
        A = rn(K, M) if a_mn else rn(M, K)
        B = rn(K, N) if b_mn else rn(N, K)
        bias = rn(N) if kw.get("bias_mode") == 1 else (rn(M) if kw.get("bias_mode") == 2 else None)
        res = rn(M, N) if kw.pop("use_res", False) else None
        if prec == BF16:
            Ab, Bb = A.bfloat16(), B.bfloat16()
            Ar, Br = Ab.double(), Bb.double()
        else:
            Ab, Bb, Ar, Br = A, B, A.double(), B.double()
        ref = (Ar.t() if a_mn else Ar) @ (Br if b_mn else Br.t())
        if bias is not None:
            ref = ref + (bias.double()[None, :] if kw["bias_mode"] == 1 else bias.double()[:, None])
        if kw.get("act") == 1:   # bf16 kernels: tanh form with the fitted cubic (csrc/common.cuh); fp32 mode: exact erf
            ref = 0.5 * ref * (1 + torch.tanh(ref * (0.80015707848 + 0.03470089342 * ref * ref))) if prec == BF16 else O.gelu_erf(ref)
        if kw.get("act") == 2:
            ref = torch.relu(ref)
        if res is not None:
            ref = ref + res.double()
        out = ops.gemm(prec, Ab, a_mn, Bb, b_mn, M, N, K, bias=bias, residual=res, **kw)
        torch.cuda.synchronize()
        return rel(out.float(), ref)

    if g == "gemm_f32":
        for (M, N, K, am, bm) in [(64, 64, 32, 0, 0), (70, 50, 37, 0, 0), (70, 50, 37, 1, 0), (70, 50, 37, 0, 1), (130, 65, 200, 1, 1)]:
            ok &= report(f"simt {M}x{N}x{K} a_mn={am} b_mn={bm}", gemm_case(FP32, M, N, K, am, bm, bias_mode=1, act=1, use_res=True), 1e-5)
    elif g == "gemm_bf16_kk":
        for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 384, 512), (200, 136, 328), (16384, 128, 3136)]:
            ok &= report(f"umma KK {M}x{N}x{K}", gemm_case(BF16, M, N, K, 0, 0), 1e-5)
        ok &= report("umma KK epilogue bias+gelu+res", gemm_case(BF16, 200, 136, 328, 0, 0, bias_mode=1, act=1, use_res=True), 1e-3)
        ok &= report("umma KK bf16 out", gemm_case(BF16, 200, 136, 328, 0, 0, bias_mode=2, out_bf16=True), 5e-3)
        ok &= report("umma KK splitk=3", gemm_case(BF16, 200, 136, 1000, 0, 0, splitk=3, bias_mode=1), 1e-5)
    elif g == "gemm_bf16_mn":
        for (am, bm) in [(0, 1), (1, 0), (1, 1)]:
            for (M, N, K) in [(128, 128, 64), (128, 128, 256), (200, 136, 328)]:
                ok &= report(f"umma a_mn={am} b_mn={bm} {M}x{N}x{K}", gemm_case(BF16, M, N, K, am, bm), 1e-5)
        ok &= report("umma wgrad-like TN 128x3072x16384 splitk=6", gemm_case(BF16, 128, 3072, 16384, 1, 1, splitk=6), 1e-5)
    elif g == "gemm_bf16_wide":
        # shapes that take the persistent 128 x 256 kernel (N > 128, >= 64 tiles): every operand layout, ragged M / N / K,
        # epilogue variants, split-K, several tiles per CTA (double-buffered accumulators), batches
        for (am, bm) in [(0, 0), (0, 1), (1, 0), (1, 1)]:
            ok &= report(f"umma2 a_mn={am} b_mn={bm} 4104x648x328", gemm_case(BF16, 4104, 648, 328, am, bm), 1e-5)
        ok &= report("umma2 KK 12544x3072x768 (13 tiles per CTA)", gemm_case(BF16, 12544, 3072, 768, 0, 0), 1e-5)
        # CTA pairs (cta_group::2, >= 74 tiles of 256 x 256): every layout, a last pair tile whose second CTA is all out of range
        for (am, bm) in [(0, 0), (0, 1), (1, 0), (1, 1)]:
            ok &= report(f"umma2c a_mn={am} b_mn={bm} 9576x648x328", gemm_case(BF16, 9576, 648, 328, am, bm), 1e-5)
        ok &= report("umma2c KK epilogue bias+gelu+res fp32 out", gemm_case(BF16, 9576, 648, 328, 0, 0, bias_mode=1, act=1, use_res=True), 1e-3)
        ok &= report("umma2c KK bf16 out row bias, long K (per-thread stores)", gemm_case(BF16, 9576, 648, 2000, 0, 0, bias_mode=2, out_bf16=True), 5e-3)
        ok &= report("umma2c KK splitk=2", gemm_case(BF16, 9576, 648, 1000, 0, 0, splitk=2, bias_mode=1), 1e-5)
        ok &= report("umma2 KK epilogue bias+gelu+res", gemm_case(BF16, 4104, 648, 328, 0, 0, bias_mode=1, act=1, use_res=True), 1e-3)
        ok &= report("umma2 KK bf16 out row bias", gemm_case(BF16, 4104, 648, 328, 0, 0, bias_mode=2, out_bf16=True), 5e-3)
        ok &= report("umma2 KK splitk=3", gemm_case(BF16, 4104, 648, 1000, 0, 0, splitk=3, bias_mode=1), 1e-5)
        ok &= report("umma2 TN wgrad-like 768x3072x12544 splitk=2", gemm_case(BF16, 768, 3072, 12544, 1, 1, splitk=2), 1e-5)
        Bt, T, N, D = 70, 384, 196, 768
        Wp = torch.zeros(T, 200, device=dev, dtype=torch.bfloat16)      # leading dimension padded to 8 elements, as the ABI does
        Wp[:, :N] = rn(T, N).bfloat16()
        W = Wp[:, :N]
        X = rn(Bt * N, D).bfloat16()
        bias = rn(T)
        out = ops.gemm(BF16, W, False, X, True, T, D, N, batch=Bt, a_batch_rows=0, b_batch_rows=N, bias=bias, bias_mode=2)
        ref = torch.einsum("tn,bnd->btd", W.double(), X.double().view(Bt, N, D)) + bias.double()[None, :, None]
        ok &= report("umma2 batched shared-A MN-major-B (token mixing of the Scaled config)", rel(out, ref), 1e-5)
    elif g == "gemm_bf16_batch":
        # token-mix shaped: shared A [T,N] (K-major), per-sample B = Xn[b] [N, D] MN-major
        Bt, T, N, D = 5, 200, 72, 136
        W = rn(T, N).bfloat16()
        X = rn(Bt * N, D).bfloat16()
        bias = rn(T)
        out = ops.gemm(BF16, W, False, X, True, T, D, N, batch=Bt, a_batch_rows=0, b_batch_rows=N, bias=bias, bias_mode=2)
        ref = torch.einsum("tn,bnd->btd", W.double(), X.double().view(Bt, N, D)) + bias.double()[None, :, None]
        ok &= report("umma batched shared-A MN-major-B", rel(out, ref), 1e-5)
    elif g == "rowops":
        x, w, b = rn(37, 5, 48), 1 + 0.1 * rn(48), 0.1 * rn(48)
        y = ops.layernorm_fwd(x, w, b)
        ok &= report("layernorm_fwd", rel(y, O.layer_norm(x.double(), w.double(), b.double())), 1e-6)
        xd = x.double().requires_grad_(True); wd = w.double().requires_grad_(True); bd = b.double().requires_grad_(True)
        dy = rn(37, 5, 48)
        O.layer_norm(xd, wd, bd).backward(dy.double())
        dx, dw, db = ops.layernorm_bwd(dy, x, w)
        ok &= report("layernorm_bwd dx", rel(dx, xd.grad), 1e-5)
        ok &= report("layernorm_bwd dw", rel(dw, wd.grad), 1e-5)
        ok &= report("layernorm_bwd db", rel(db, bd.grad), 1e-5)
        img = rn(3, 2, 8, 12)
        cols = ops.patch_gather(img, 4)
        wconv = rn(7, 2, 4, 4)
        ref = O.patch_embed(img.double(), wconv.double(), torch.zeros(7, device=dev, dtype=torch.float64))
        ok &= report("patch_gather", rel(cols.double() @ wconv.double().reshape(7, -1).t(), ref), 1e-12)
        a, c = rn(4, 3, 16), rn(4, 5, 16)
        cat = ops.concat_tokens([a, c])
        ok &= report("concat_tokens", rel(cat, torch.cat([a, c], 1)) + 0.0, 1e-12)
        sp = ops.split_tokens(cat, [3, 5])
        ok &= report("split_tokens", rel(sp[1], c) + rel(sp[0], a), 1e-12)
        ok &= report("add", rel(ops.add(a, a), 2 * a), 1e-12)
        wb = ops.cast_bf16(rn(5, 13), 16)
        ok &= report("cast_bf16 pad", float(wb[:, 13:].float().abs().sum()), 1e-12)
    elif g == "token_mix":
        for (B, N, D, T, prec, tol) in [(9, 4, 128, 32, FP32, 2e-6), (9, 4, 128, 32, BF16, 1e-2), (5, 25, 64, 8, FP32, 2e-6),
                                        (3, 40, 48, 16, FP32, 2e-6), (2, 196, 96, 64, FP32, 2e-6), (70, 8, 128, 32, BF16, 1e-2),
                                        (5, 4, 32, 8, BF16, 1e-2), (11, 8, 64, 16, BF16, 1e-2), (3, 12, 64, 32, BF16, 1e-2),
                                        (6, 4, 256, 32, BF16, 1e-2), (4, 25, 64, 8, BF16, 1e-2), (2100, 4, 128, 32, BF16, 1e-2),
                                        (7, 24, 64, 16, BF16, 1e-2), (130, 32, 32, 32, BF16, 1e-2), (5, 17, 64, 20, BF16, 1e-2)]:
            x = rn(B, N, D)
            p = dict(ln_w=1 + 0.1 * rn(D), ln_b=0.1 * rn(D), w1=rn(T, N) / N ** 0.5, b1=0.1 * rn(T), w2=rn(N, T) / T ** 0.5, b2=0.1 * rn(N))
            pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
            xd = x.double().requires_grad_(True)
            xn = O.layer_norm(xd, pd["ln_w"], pd["ln_b"])
            h = O.gelu_erf(torch.einsum("tn,bnd->btd", pd["w1"], xn) + pd["b1"][None, :, None])
            ur = xd + torch.einsum("nt,btd->bnd", pd["w2"], h) + pd["b2"][None, :, None]
            u = ops.token_mix_fwd(x, p["ln_w"], p["ln_b"], p["w1"], p["b1"], p["w2"], p["b2"], prec)
            tag = f"token_mix B{B} N{N} D{D} T{T} p{prec}"
            ok &= report(tag + " fwd", rel(u, ur), tol)
            ok &= report(tag + " fwd (branch only)", rel(u - x, ur.detach() - xd.detach()), tol)
            du = rn(B, N, D)
            ur.backward(du.double())
            outs = ops.token_mix_bwd(du, x, p["ln_w"], p["ln_b"], p["w1"], p["b1"], p["w2"], prec)
            names = ["dx", "ln_w", "ln_b", "w1", "b1", "w2", "b2"]
            refs = [xd.grad] + [pd[k].grad for k in names[1:]]
            for nme, o, r in zip(names, outs, refs):
                ok &= report(tag + " bwd " + nme, rel(o, r), tol * (2 if prec == BF16 else 20))
    elif g in ("chain_f32", "chain_fwd", "chain_bwd", "chain_unfused"):
        if g == "chain_f32":
            cases = [(300, 48, 70, FP32, 3e-6), (1000, 128, 3078, FP32, 3e-6)]
        elif g == "chain_unfused":
            cases = [(300, 384, 520, BF16, 1e-2)]
        else:
            cases = [(128, 128, 64, BF16, 1e-2), (128, 128, 256, BF16, 1e-2), (300, 128, 3078, BF16, 1e-2),
                     (16384, 128, 3072, BF16, 1e-2), (300, 64, 200, BF16, 1e-2), (300, 32, 256, BF16, 1e-2)]
            if g == "chain_fwd":
                cases.append((300, 256, 512, BF16, 1e-2))
        for (M, D, Cc, prec, tol) in cases:
            u = rn(M, D)
            p = dict(ln_w=1 + 0.1 * rn(D), ln_b=0.1 * rn(D), w1=rn(Cc, D) / D ** 0.5, b1=0.1 * rn(Cc), w2=rn(D, Cc) / Cc ** 0.5, b2=0.1 * rn(D))
            w1b = w2b = None
            if prec == BF16:
                w1b, w2b = ops.cast_bf16(p["w1"], D), ops.cast_bf16(p["w2"], (Cc + 7) // 8 * 8)
            pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
            ud = u.double().requires_grad_(True)
            yr = ud + O.gelu_erf(O.layer_norm(ud, pd["ln_w"], pd["ln_b"]) @ pd["w1"].t() + pd["b1"]) @ pd["w2"].t() + pd["b2"]
            tag = f"channel_mix M{M} D{D} C{Cc} p{prec}"
            if g != "chain_bwd":
                y = ops.channel_mix_fwd(u, p["ln_w"], p["ln_b"], p["w1"], p["b1"], p["w2"], p["b2"], w1b, w2b, prec)
                torch.cuda.synchronize()
                ok &= report(tag + " fwd", rel(y, yr), tol)
                ok &= report(tag + " fwd (branch only)", rel(y - u, yr.detach() - ud.detach()), tol)
            if g != "chain_fwd":
                dy = rn(M, D)
                yr.backward(dy.double())
                outs = ops.channel_mix_bwd(dy, u, p["ln_w"], p["ln_b"], p["w1"], p["b1"], p["w2"], w1b, w2b, prec)
                torch.cuda.synchronize()
                names = ["du", "ln_w", "ln_b", "w1", "b1", "w2", "b2"]
                refs = [ud.grad] + [pd[k].grad for k in names[1:]]
                for nme, o, r in zip(names, outs, refs):
                    ok &= report(tag + " bwd " + nme, rel(o, r), tol * (2 if prec == BF16 else 20))
    elif g == "linear":
        for prec, tol in [(FP32, 3e-6), (BF16, 1e-2)]:
            x, w, b = rn(50, 6, 40), rn(24, 40) / 40 ** 0.5, 0.1 * rn(24)
            wb = ops.cast_bf16(w, 40) if prec == BF16 else None
            for act in (0, 2):
                xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
                yr = xd @ wd.t() + bd
                yr = torch.relu(yr) if act == 2 else yr
                y = ops.linear_fwd(x, w, wb, b, act, prec)
                ok &= report(f"linear fwd act{act} p{prec}", rel(y, yr), tol)
                dy = rn(50, 6, 24)
                yr.backward(dy.double())
                dx, dw, db = ops.linear_bwd(dy, x, y, w, wb, act, True, prec)
                if act == 0 or prec == FP32:
                    ok &= report(f"linear bwd dx act{act} p{prec}", rel(dx, xd.grad), tol * 3)
                    ok &= report(f"linear bwd dw act{act} p{prec}", rel(dw, wd.grad), tol * 3)
                    ok &= report(f"linear bwd db act{act} p{prec}", rel(db, bd.grad), tol * 3)
    elif g == "heads":
        for kind, B, K in ((0, 37, 10), (1, 37, 23), (0, 1000, 10), (1, 300, 12)):   # K <= 16: vectorised kernels; 23: generic
            toks = [rn(B, 4, 128), rn(B, 1, 64), rn(B, 8, 128)]
            ws = [rn(K, 128) / 11, rn(K, 64) / 8, rn(K, 128) / 11]
            bs = [0.1 * rn(K) for _ in range(3)]
            hw = [0.7, 1.1, 1.3]
            pw = (1 + 5 * torch.rand(K, device=dev)) if kind else None
            labels = torch.randint(0, K, (B,), device=dev) if kind == 0 else (torch.rand(B, K, device=dev) < 0.2).float()
            td = [t.double().requires_grad_(True) for t in toks]
            wd = [t.double().requires_grad_(True) for t in ws]
            bd = [t.double().requires_grad_(True) for t in bs]
            lg = [O.pooled_linear(t, w, b) for t, w, b in zip(td, wd, bd)]
            Ls = [O.cross_entropy(l, labels) if kind == 0 else O.bce_pos_weight(l, labels.double(), pw.double()) for l in lg]
            tot = sum(h * L for h, L in zip(hw, Ls))
            losses, logits, preds = ops.heads_loss_fwd(toks, ws, bs, labels, pw, hw, kind)
            ok &= report(f"heads kind{kind} B{B} K{K} logits", rel(logits, torch.stack(lg)), 1e-5)
            ok &= report(f"heads kind{kind} B{B} K{K} losses", rel(losses, torch.stack([tot] + Ls)), 1e-5)
            pr = torch.stack([l.argmax(1) for l in lg]) if kind == 0 else torch.stack([(l > 0).long() for l in lg])
            ok &= report(f"heads kind{kind} B{B} K{K} preds", float((preds != pr).sum()), 0.5)
            tot.backward()
            dt, dw, db = ops.heads_loss_bwd(toks, ws, bs, labels, pw, hw, kind, logits, 1.0, None)
            for i in range(3):
                ok &= report(f"heads kind{kind} B{B} K{K} dtok{i}", rel(dt[i], td[i].grad), 1e-5)
                ok &= report(f"heads kind{kind} B{B} K{K} dw{i}", rel(dw[i], wd[i].grad), 1e-5)
                ok &= report(f"heads kind{kind} B{B} K{K} db{i}", rel(db[i], bd[i].grad), 1e-5)
    elif g == "patch_embed":
        # (B, cin, H, W, P, D): fused gather-GEMM shapes (P % 8 == 0), ragged M / D / K tiles, and the fallback (P = 14, fp32)
        for (B, cin, H, W, P, D) in [(64, 1, 112, 112, 56, 128), (37, 3, 32, 48, 16, 200), (5, 2, 16, 24, 8, 72),
                                     (300, 1, 112, 112, 56, 128), (9, 1, 28, 28, 14, 128)]:
            for prec, in_bf16 in ((BF16, False), (BF16, True), (FP32, False)):
                img = rn(B, cin, H, W)
                w, b = rn(D, cin, P, P) / (cin * P * P) ** 0.5, 0.1 * rn(D)
                dy = rn(B, (H // P) * (W // P), D)
                imgr = img.bfloat16().double() if prec == BF16 else img.double()
                wr = (w.bfloat16().double() if prec == BF16 else w.double()).requires_grad_(True)
                bd = b.double().requires_grad_(True)
                ref = O.patch_embed(imgr, wr, bd)
                ref.backward((dy.bfloat16() if prec == BF16 else dy).double())
                wb = ops.cast_bf16(w.reshape(D, -1), (cin * P * P + 7) // 8 * 8) if prec == BF16 else None
                x = img.bfloat16() if in_bf16 else img
                y = ops.patch_embed_fwd(x, w, wb, b, P, prec)
                dw, db = ops.patch_embed_bwd(dy, x, w, P, True, prec)
                torch.cuda.synchronize()
                tag = f"patch_embed B{B} c{cin} {H}x{W} P{P} D{D} p{prec} bf16in={int(in_bf16)}"
                ok &= report(tag + " fwd", rel(y, ref), 1e-5)
                ok &= report(tag + " dw", rel(dw, wr.grad), 2e-5)
                ok &= report(tag + " db", rel(db, dy.double().sum((0, 1))), 2e-6)
    elif g == "adam":
        n = 100003
        p0, gr = rn(n), rn(n)
        p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        pr = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([pr], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
        state = torch.tensor([1e-2, 0.0], device=dev)
        p2, m2, v2 = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        for step in range(1, 4):
            pr.grad = gr.clone()
            opt.step()
            ops.adam_step(p, gr, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.01, step, 1.0, None)
            ops.adam_step(p2, gr, m2, v2, 0.0, 0.9, 0.999, 1e-8, 0.01, 0, 1.0, state)
        ok &= report("adam 3 steps (host scalars)", rel(p, pr.detach()), 1e-6)
        ok &= report("adam 3 steps (device state)", rel(p2, pr.detach()), 1e-6)
    torch.cuda.synchronize()
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--group":
        sys.exit(0 if run_group(sys.argv[2]) else 1)
    groups = sys.argv[1:] or GROUPS
    failed = 0
    for g in groups:
        print(f"== {g}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], timeout=240)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
        print(f"== {g}: rc={rc} ({time.time() - t0:.1f}s)", flush=True)
        failed += rc != 0
    print(f"SELF-CHECK: {len(groups) - failed}/{len(groups)} groups passed")
    sys.exit(failed)
