"""Shared helpers for the parity tests (the oracle is the checker, never the thing under test)."""
import torch


def rel_max_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """SURVEY.md 8d parity metric: max|cand - ref| / max|ref| over the whole tensor."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def describe_mismatch(name: str, got: torch.Tensor, ref: torch.Tensor, tol: float) -> str:
    """Human-readable error map so that ONE gpu run tells where a kernel went wrong."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if got.dim() == 1:
        got, ref = got[None], ref[None]
    got2, ref2 = got.reshape(-1, got.shape[-1]), ref.reshape(-1, ref.shape[-1])
    err = (got2 - ref2).abs()
    scale = ref2.abs().max().clamp_min(1e-30)
    bad = err > tol * scale
    lines = [f"[{name}] shape={tuple(got.shape)} rel_max_err={float(err.max() / scale):.4g} tol={tol} "
             f"bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(got2).sum())} "
             f"got_absmax={float(got2.abs().max()):.4g} ref_absmax={float(scale):.4g}"]
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        lines.append(f"  bad rows: n={len(rows)} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        lines.append(f"  bad cols: n={len(cols)} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        R, Cc = got2.shape
        rb, cb = max(1, min(R, 256) // 16), max(1, min(Cc, 256) // 16)
        sub = bad[:rb * 16, :cb * 16].float().reshape(16, rb, 16, cb).mean(dim=(1, 3)) if R >= 16 and Cc >= 16 else None
        if sub is not None:
            lines.append(f"  bad-fraction map of the first {rb*16}x{cb*16} block (16x16 cells, rows={rb}/cell, cols={cb}/cell):")
            for r in range(16):
                lines.append("   " + " ".join(f"{int(9.99 * v)}" for v in sub[r].tolist()))
        i = int(err.argmax())
        r, c = divmod(i, got2.shape[1])
        lines.append(f"  worst at ({r},{c}): got={float(got2[r, c]):.6g} ref={float(ref2[r, c]):.6g}")
        lines.append(f"  got[0,:8]={[round(float(v), 4) for v in got2[0, :8]]}")
        lines.append(f"  ref[0,:8]={[round(float(v), 4) for v in ref2[0, :8]]}")
    return "\n".join(lines)


def parity_report(name: str, got: torch.Tensor, ref: torch.Tensor, tol: float = 2e-2) -> dict:
    """The two stricter readings of "max relative error <= 2e-2" that SURVEY.md 8d asks to report next to the normalised max:
    element-wise pass-rate of |d| <= tol |ref| + tol rms(ref), and the minimum per-row cosine similarity."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    g2, r2 = got.reshape(-1, got.shape[-1]), ref.reshape(-1, ref.shape[-1])
    rms = float(r2.pow(2).mean().sqrt())
    ok = (g2 - r2).abs() <= tol * r2.abs() + tol * rms
    norm = g2.norm(dim=1) * r2.norm(dim=1)
    live = norm > 0
    cos = ((g2 * r2).sum(1)[live] / norm[live]) if bool(live.any()) else torch.ones(1)
    rep = {"pass_rate": float(ok.float().mean()), "min_row_cosine": float(cos.min()), "rel_max_err": rel_max_err(got, ref)}
    print(f"[{name}] elementwise pass-rate {rep['pass_rate']:.6f} (|d| <= {tol}|ref| + {tol} rms), "
          f"min row cosine {rep['min_row_cosine']:.6f}, normalised max err {rep['rel_max_err']:.4g}")
    return rep


def assert_parity(name: str, got: torch.Tensor, ref: torch.Tensor, tol: float = 2e-2, min_pass_rate: float = 0.999) -> dict:
    """Path-level bar: normalised max error <= tol AND element-wise pass-rate >= 0.999; the row cosine is reported."""
    assert_close(name, got, ref, tol)
    rep = parity_report(name, got, ref, tol)
    assert rep["pass_rate"] >= min_pass_rate, f"{name}: element-wise pass-rate {rep['pass_rate']:.6f} < {min_pass_rate}"
    return rep


def assert_close(name: str, got: torch.Tensor, ref: torch.Tensor, tol: float) -> None:
    e = rel_max_err(got, ref)
    ok = e <= tol and not bool(torch.isnan(got.float()).any())
    msg = describe_mismatch(name, got, ref, tol)
    print(msg.splitlines()[0])
    assert ok, "\n" + msg
