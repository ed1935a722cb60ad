"""GPU parity of the encoder training path (--train-bio): forward with the tape == inference forward, and
the gradient of EVERY encoder parameter against fp32 autograd of the oracle (the reference's HF modules restated)."""
import pytest
import torch

from oracle import synth
from oracle.esm_oracle import SPECS, esm_encoder_forward, init_encoder_weights, init_projector
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2


def _ids(spec, K, valids, seed):
    g = torch.Generator().manual_seed(seed)
    rows = [synth.protein_ids(g, K, v) if spec.vocab_size == 33 else synth.nucleotide_ids(g, K, v, spec.vocab_size)
            for v in valids]
    ids = torch.stack(rows)
    if spec.token_dropout:
        ids[0, 4] = spec.mask_token_id                      # token-dropout rescale path
    return ids


@pytest.mark.parametrize("spec_name,K,valids", [
    ("tiny_esm2", 150, [150, 77, 130]),                    # rotary, erf-GELU, token dropout, K not a multiple of 128
    ("tiny_ntv2", 140, [140, 9]),                          # gated SiLU, no biases in the FFN
    ("tiny_ntv1", 96, [96, 40, 96]),                       # learned absolute positions
    ("esm2_t6_8m", 256, [256, 100]),                       # head_dim 16
    ("nt_v2_50m", 130, [130, 64]),                         # head_dim 32
])
def test_encoder_backward_vs_oracle_autograd(spec_name, K, valids):
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    from molly_b200 import train
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 7)
    proj = init_projector(spec.hidden_size, 64, 8)
    ids = _ids(spec, K, valids, 9)
    g = torch.Generator().manual_seed(10)
    d_out = (torch.randn(len(valids) * K, spec.hidden_size, generator=g) * 0.1).to(torch.bfloat16)
    # oracle: fp32 autograd on the CPU
    Wg = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in W.items()}
    ref_out = esm_encoder_forward(spec, Wg, ids)
    (ref_out.reshape(-1, spec.hidden_size) * d_out.float()).sum().backward()
    # the reference's own execution: stock HF modules in bf16 on the GPU (what --train-bio trains with), same d_out
    from oracle.ref_import import build_hf_encoder
    hf = build_hf_encoder(spec, W).to(DEV).to(torch.bfloat16).train(False)
    for prm in hf.parameters():
        prm.requires_grad_(True)
    hf_out = hf(input_ids=ids.to(DEV), attention_mask=(ids != 1).to(DEV), output_hidden_states=True).hidden_states[-1]
    (hf_out.reshape(-1, spec.hidden_size).float() * d_out.to(DEV).float()).sum().backward()
    hf_grads = {k: v.grad.float().cpu() for k, v in hf.named_parameters() if v.grad is not None}
    # candidate
    enc = PackedEncoder(EncoderConfig.from_mapping(spec.as_dict()), W, proj, K, torch.device(DEV))
    try:
        out, tape = train.encoder_forward_train(enc, ids.to(DEV))
        assert_close(f"{spec_name} train-mode forward", out.float().cpu(), ref_out.detach().reshape(-1, spec.hidden_size), TOL)
        grads = train.encoder_backward(enc, tape, d_out.to(DEV))
        checked, worst_fro, worst_ratio = 0, 0.0, 0.0
        for name, ref in Wg.items():
            if ref.grad is None or name.startswith("lm_head") or "contact_head" in name or "inv_freq" in name:
                continue
            assert name in grads, f"no gradient produced for {name}"
            got = grads[name].cpu()
            if float(ref.grad.abs().max()) == 0.0:
                assert float(got.abs().max()) == 0.0, name
            else:
                # gradients are sums of many bf16-rounded products (dS, P, activations): the aggregate error is the bar
                # (relative Frobenius <= 2e-2); single elements of small-magnitude tensors (q / k weights) may sit at 3-5e-2
                scale_n, scale_m = ref.grad.norm(), ref.grad.abs().max()
                if name.endswith("key.bias"):
                    # softmax is invariant to a constant added to every key, so d(key bias) is a near-total cancellation
                    # (exactly 0 without rotary): judge it on the scale of the query-bias gradient of the same layer
                    q = Wg[name.replace("key.bias", "query.bias")].grad
                    scale_n, scale_m = max(scale_n, q.norm()), max(scale_m, q.abs().max())
                fro = float((got - ref.grad).norm() / scale_n)
                worst = float((got - ref.grad).abs().max() / scale_m)
                # the bar: 2e-2, or -- where bf16 itself cannot do better (d(q), d(k) pass through dS = P * (dP - delta), a
                # cancellation of bf16-rounded terms) -- the error of the reference's own bf16 backward on the same weights
                hf_fro = float((hf_grads[name] - ref.grad).norm() / scale_n) if name in hf_grads else 0.0
                tol = max(TOL, 2.0 * hf_fro)
                assert fro <= tol, (f"{spec_name} d {name}: rel Frobenius {fro:.4f} (reference's bf16 backward: {hf_fro:.4f}), "
                                    f"normalised max {worst:.4f}")
                worst_ratio = max(worst_ratio, fro / max(hf_fro, 1e-9))
                worst_fro = max(worst_fro, fro)
            checked += 1
        assert checked >= 12 * spec.num_hidden_layers
        print(f"[{spec_name}] {checked} parameter gradients checked, worst relative Frobenius error {worst_fro:.4f}, "
              f"at most {worst_ratio:.2f}x the error of the reference's bf16 backward")
    finally:
        enc.close()


@pytest.mark.parametrize("enc_name", ["esm2_t33_650m", "nt_v2_500m"])
def test_encoder_backward_full_size_models(enc_name):
    """The two BASELINE configs[1] encoders at full depth and width (33 x 1280 / 29 x 1024, head_dim 64), 2 x 1024 tokens:
    gradients against autograd of the fp32 oracle run on the GPU (TF32 off).  Early-layer q / k gradients of a deep bf16
    network are noisy in any implementation, so the aggregate over ALL parameters is the criterion here: relative error of
    the concatenated gradient <= 2e-2 and every tensor <= 10 %."""
    import bench
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    from molly_b200 import train
    torch.backends.cuda.matmul.allow_tf32 = False
    e = bench.ENC[enc_name]
    spec = SPECS[enc_name]
    dev = torch.device(DEV, 0)
    sd = {k: v.to(torch.bfloat16).float() for k, v in bench.gpu_state_dict(e, dev, 123).items()}
    K, valids = 1024, [1024, 700]
    ids = _ids(spec, K, valids, 31).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(32)
    d_out = (torch.randn(len(valids) * K, spec.hidden_size, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    Wg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_out = esm_encoder_forward(spec, Wg, ids)
    (ref_out.reshape(-1, spec.hidden_size) * d_out.float()).sum().backward()
    proj = {"weight": torch.zeros(64, spec.hidden_size), "bias": torch.zeros(64)}
    enc = PackedEncoder(EncoderConfig.from_mapping(dict(e, name=enc_name)), sd, proj, K, dev)
    try:
        out, tape = train.encoder_forward_train(enc, ids)
        assert_close(f"{enc_name} train-mode forward", out.float(), ref_out.detach().reshape(-1, spec.hidden_size), TOL)
        grads = train.encoder_backward(enc, tape, d_out)
        num = den = 0.0
        worst, worst_name = 0.0, ""
        for name, ref in Wg.items():
            if ref.grad is None:
                continue
            diff = float((grads[name] - ref.grad).norm())
            num += diff ** 2
            den += float(ref.grad.norm()) ** 2
            scale = float(ref.grad.norm())
            if ".query." in name or ".key." in name:
                # with near-uniform attention (random-init weights) d(q), d(k) are near-total cancellations, orders of
                # magnitude below d(v): judge them on the scale of the value gradient of the same layer
                v_name = name.replace(".query.", ".value.").replace(".key.", ".value.")
                scale = max(scale, float(Wg[v_name].grad.norm()))
            rel = diff / max(scale, 1e-30)
            if rel > worst:
                worst, worst_name = rel, name
        total = (num / den) ** 0.5
        print(f"[{enc_name}] all-parameter gradient: relative error {total:.4f}; worst tensor {worst_name} {worst:.4f}")
        assert total <= TOL and worst <= 0.10, (total, worst, worst_name)
    finally:
        enc.close()


def test_native_train_abi_contract():
    """``molly_encoder_train_sizes`` / ``molly_encoder_grad_layout`` / ``molly_encode_train_fwd|bwd`` (include/molly_b200.h): the
    flat gradient layout covers every encoder parameter exactly once, a tape / workspace that is too small and a layer range
    outside the encoder are refused with a status and a message (no kernel is launched), and a backward issued layer range by
    layer range (what the reducer path does) equals the one-call backward."""
    import ctypes as C
    from molly_b200 import _lib, ops, train
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    for spec_name in ("tiny_esm2", "tiny_ntv2", "tiny_ntv1"):
        spec = SPECS[spec_name]
        W = init_encoder_weights(spec, 3)
        enc = PackedEncoder(EncoderConfig.from_mapping(spec.as_dict()), W, init_projector(spec.hidden_size, 64, 4), 64,
                            torch.device(DEV))
        try:
            plan = train.grad_plan(enc)
            n_params = sum(v.numel() for k, v in W.items() if v.dtype.is_floating_point and not k.endswith("inv_freq")
                           and "contact_head" not in k and "lm_head" not in k and "pooler" not in k)
            assert plan.total == n_params, (spec_name, plan.total, n_params)
            spans = []
            for i in range(plan.L):
                for name, slot, sub, shape in plan._layer_entries(i):
                    n = 1
                    for v in shape:
                        n *= v
                    assert tuple(W[name].shape) == shape or plan.glu, name
                    spans.append((i * plan.group + plan.off[slot] + sub, n))
            for name, slot, shape in plan._tail_entries():
                n = 1
                for v in shape:
                    n *= v
                spans.append((plan.tail[slot], n))
            spans.sort()
            assert spans[0][0] == 0 and all(a + n == b for (a, n), (b, _) in zip(spans, spans[1:]))      # a partition
            assert spans[-1][0] + spans[-1][1] == plan.total
            K, n_seq = 64, 2
            ids = _ids(spec, K, [64, 30], 5).to(DEV)
            tape_b, ws_b, gf = train._sizes(enc, n_seq, K, False)
            assert gf == plan.total and tape_b > 0 and ws_b > 0
            assert train._sizes(enc, n_seq, K, True)[0] < tape_b                  # the recompute tape keeps one activation slot
            lib = _lib.load()
            out = torch.empty(n_seq * K, spec.hidden_size, dtype=torch.bfloat16, device=DEV)
            small = train._device_buffer(tape_b - 1024, torch.device(DEV))
            rc = lib.molly_encode_train_fwd(enc.handle, ids.data_ptr(), n_seq, K, out.data_ptr(), small.data_ptr(), tape_b - 1024, 0,
                                            ops.error_flag(torch.device(DEV)).data_ptr(), ops._stream(torch.device(DEV)))
            assert rc == 4 and b"tape" in lib.molly_last_error()                   # MOLLY_STATUS_WORKSPACE
            out, tape = train.encoder_forward_train(enc, ids)
            g = torch.Generator().manual_seed(1)
            d_out = (torch.randn(n_seq * K, spec.hidden_size, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
            flat = torch.zeros(plan.total, dtype=torch.float32, device=DEV)
            ws = train._device_buffer(ws_b, torch.device(DEV))

            def bwd(lb, le, ws_bytes=ws_b):
                return lib.molly_encode_train_bwd(enc.handle, n_seq, K, tape.buf.data_ptr(), tape_b, 0, d_out.data_ptr(),
                                                  flat.data_ptr(), ws.data_ptr(), ws_bytes, lb, le, ops._stream(torch.device(DEV)))
            assert bwd(plan.L + 1, 0) == 1 and bwd(0, 1) == 1 and bwd(plan.L, -1) == 1      # MOLLY_STATUS_INVALID
            assert bwd(plan.L, 0, ws_b - 1024) == 4
            whole = train.encoder_backward(enc, tape, d_out)
            assert bwd(plan.L, plan.L - 1) == 0                                              # range by range, top down
            for i in range(plan.L - 2, -1, -1):
                assert bwd(i, i) == 0
            torch.cuda.synchronize()
            for i in range(plan.L):
                part = plan.finalize_(plan.layer_views(i, flat[i * plan.group:(i + 1) * plan.group]))
                for name, t in part.items():
                    assert_close(f"{spec_name} {name} (layer ranges vs one call)", t.float().cpu(), whole[name].float().cpu(), 1e-3)
        finally:
            enc.close()


@pytest.mark.parametrize("spec_name", ["tiny_esm2", "tiny_ntv2", "tiny_ntv1"])
def test_graphed_training_step_equals_eager(spec_name):
    """From the third call with a shape on, ``encoder_forward_train`` / ``encoder_backward`` replay CUDA graphs (``_GraphedStep``):
    same outputs and gradients as the eager launches (column sums and split-K partials meet in fp32 atomics / reduce-adds, so
    equal up to their order), also when the ids and d_out CHANGE between replays (the graphs read static buffers), the forward
    output of a graphed step survives until its backward, and a re-entrant forward falls back to the eager path."""
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    from molly_b200 import ops, train
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 11)
    K, valid_sets = 96, ([96, 50, 70], [33, 96, 96], [96, 96, 5])
    g = torch.Generator().manual_seed(12)
    d_outs = [(torch.randn(3 * K, spec.hidden_size, generator=g) * 0.1).to(torch.bfloat16).to(DEV) for _ in range(3)]
    idss = [_ids(spec, K, v, 20 + i).to(DEV) for i, v in enumerate(valid_sets)]

    def one(enc, i):
        out, tape = train.encoder_forward_train(enc, idss[i])
        out = out.clone()
        grads = train.encoder_backward(enc, tape, d_outs[i])
        return out, {k: v.float().clone() for k, v in grads.items()}, tape

    import os
    os.environ["MOLLY_TRAIN_GRAPH"] = "0"
    enc = PackedEncoder(EncoderConfig.from_mapping(spec.as_dict()), W, init_projector(spec.hidden_size, 64, 4), K, torch.device(DEV))
    try:
        eager = [one(enc, i)[:2] for i in range(3)]
    finally:
        enc.close()
        del os.environ["MOLLY_TRAIN_GRAPH"]
    enc = PackedEncoder(EncoderConfig.from_mapping(spec.as_dict()), W, init_projector(spec.hidden_size, 64, 4), K, torch.device(DEV))
    try:
        for _ in range(2):                                         # the first two calls with this shape: eager
            l0 = ops.kernel_launch_count()
            o0, g0, t0 = one(enc, 0)
            per_step = ops.kernel_launch_count() - l0
            assert t0.graph is None
        for i in (1, 2, 0):                                        # capture + replay, replay, replay with the first inputs again
            l0 = ops.kernel_launch_count()
            out, grads, tape = one(enc, i)
            assert tape.graph is not None and tape.graph.fwd is not None and tape.graph.bwd is not None
            assert ops.kernel_launch_count() - l0 == per_step, "replays must be counted like the eager launches"
            assert torch.equal(out, eager[i][0]), "graphed forward differs from the eager forward"
            for name, ref in eager[i][1].items():
                assert_close(f"{spec_name} graphed d {name} (inputs {i})", grads[name].cpu(), ref.cpu(), 1e-4)
        # a second forward before the first one's backward must not touch the first one's buffers
        out_a, tape_a = train.encoder_forward_train(enc, idss[1])
        out_b, tape_b = train.encoder_forward_train(enc, idss[2])
        assert tape_a.graph is not None and tape_b.graph is None
        assert torch.equal(out_a, eager[1][0]) and torch.equal(out_b, eager[2][0])
        gb = train.encoder_backward(enc, tape_b, d_outs[2])
        ga = train.encoder_backward(enc, tape_a, d_outs[1])
        for name in ("esm.encoder.layer.0.attention.self.value.weight", "esm.embeddings.word_embeddings.weight"):
            assert_close(f"{spec_name} re-entrant a {name}", ga[name].float().cpu(), eager[1][1][name].cpu(), 1e-4)
            assert_close(f"{spec_name} re-entrant b {name}", gb[name].float().cpu(), eager[2][1][name].cpu(), 1e-4)
        del tape_a, tape_b
        # a forward whose backward never comes does not block the graphs for good: it is passed over a few times (eager steps),
        # then its buffers are reclaimed, and its backward -- should it still come -- is refused
        out_c, tape_c = train.encoder_forward_train(enc, idss[0])
        assert tape_c.graph is not None
        for _ in range(train._GRAPH_ABANDON):
            _, t, = train.encoder_forward_train(enc, idss[1])
            assert t.graph is None
            train.encoder_backward(enc, t, d_outs[1])
        out_d, tape_d = train.encoder_forward_train(enc, idss[2])
        assert tape_d.graph is not None and torch.equal(out_d, eager[2][0])
        gd = train.encoder_backward(enc, tape_d, d_outs[2])
        assert_close(f"{spec_name} after reclaiming", gd["esm.embeddings.word_embeddings.weight"].float().cpu(),
                     eager[2][1]["esm.embeddings.word_embeddings.weight"].cpu(), 1e-4)
        with pytest.raises(RuntimeError, match="taken over"):
            train.encoder_backward(enc, tape_c, d_outs[0])
    finally:
        enc.close()
