#!/usr/bin/env python
""""also time the reference on GPU in bf16 -- that is the real bar for beating the library kernels".
What the reference executes for the protein modality is HF `EsmForMaskedLM` (omics_one.py:83-88) + `nn.Linear`; this
script times exactly that stack (stock transformers, random-init ESM-2 650M, bf16, eager and SDPA attention, LM head
included because the reference computes and discards it) against this repo's encoder + projector on the same ids, on
one GPU.  NT-v2's gated FFN is hub remote code (not on disk), so only the protein half is compared."""
import json
import statistics
import sys
import torch
sys.path.insert(0, ".")
import bench



def time_fn(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    from transformers import EsmConfig, EsmForMaskedLM
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS["molly_1p7b"]
    e = bench.ENC[wl["pr"]]
    n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    K, D = wl["K"], wl["D"]
    omic_ids, _ = bench.make_inputs(dict(wl, B=n_seq))
    ids = omic_ids[:, 1].contiguous().to(dev)                       # the protein sequences
    out = {"encoder": wl["pr"], "n_seq": n_seq, "K": K, "D": D, "tokens": n_seq * K}
    proj = torch.nn.Linear(e["hidden_size"], D, device=dev, dtype=torch.bfloat16)
    for impl in ("sdpa", "eager"):
        cfg = EsmConfig(vocab_size=e["vocab_size"], hidden_size=e["hidden_size"], num_hidden_layers=e["num_hidden_layers"],
                        num_attention_heads=e["num_attention_heads"], intermediate_size=e["intermediate_size"],
                        position_embedding_type="rotary", token_dropout=True, mask_token_id=32, pad_token_id=1,
                        layer_norm_eps=1e-5, emb_layer_norm_before=False, attn_implementation=impl)
        torch.manual_seed(0)
        model = EsmForMaskedLM(cfg).to(dev).to(torch.bfloat16).eval()
        mask = ids != 1

        @torch.no_grad()
        def ref_step():                                             # omics_one.py:83-91 + :92
            o = model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
            return proj(o.hidden_states[-1])

        try:
            ms = time_fn(ref_step)
            out[f"hf_{impl}_ms"] = round(ms, 3)
            out[f"hf_{impl}_tokens_per_s"] = round(n_seq * K / ms * 1e3, 1)
        except torch.OutOfMemoryError as ex:
            out[f"hf_{impl}_ms"] = f"OOM: {str(ex)[:80]}"
        del model
        torch.cuda.empty_cache()
    path = bench.build_path(wl, dev, strict=False)
    hs = torch.zeros(n_seq, K + 8, D, dtype=torch.bfloat16, device=dev)
    infos = [[{"type": "protein", "start": 2}] for _ in range(n_seq)]
    ids3 = ids.view(n_seq, 1, K)
    ms = time_fn(lambda: path.process_omic_sequences(hs, ids3, infos, dev))
    out["molly_b200_ms"] = round(ms, 3)
    out["molly_b200_tokens_per_s"] = round(n_seq * K / ms * 1e3, 1)
    path.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
