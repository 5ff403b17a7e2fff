#!/usr/bin/env python
"""Generates tests/golden/hf_t5_tiny.npz and tests/golden/hf_t5_full.npz: outputs of the PUBLIC T5 definition
(HF transformers ``T5EncoderModel``, CPU fp32) plus a torch Conv1d 3Di head, on synthetic ProstT5-shaped weights:
the TINY config (unicore_b200.synth seed 7) and the FULL ProstT5 config (24 layers, d 1024, 32 heads, d_ff 16384,
seed 1 = the weights of bench.py and of the full-size GPU tests).  The oracles (oracle/prostt5_oracle.py and
oracle/prostt5_oracle.c, rounding policy "none") must reproduce these to fp32 round-off; that is the pin of the
oracles' encoder/head arithmetic (the reference itself pins nothing, see the oracle headers).

The full-size file holds, for one 350-aa (config 2, sequence 0) and one 1,200-aa sequence: the logits of every
residue and the final hidden state of every 8th token row (fp32) - 24 layers deep, so the fp16 policy of the CUDA
path (saturation, fp16 P, ex2.approx) is checked against an independent implementation at depth.

Run here (needs transformers + torch, CPU only):   python tests/golden/make_hf_golden.py [tiny|full|all]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from transformers import T5Config, T5EncoderModel  # noqa: E402

from unicore_b200 import prostt5_spec as spec, synth  # noqa: E402

SEED = 7
SEQS = [b"MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQAPILSRVGDGTQDNLSGAEKAVQVKVKALPDAQFEVVHSLAKWKRQTLGQHDFSAGEGLYTHMKALRPDEDRLSPLHSVYVDQWDWERVMGDGERQFSTLKSTVEAIWAGIKATEAAVSEEFGLAPFLPDQIHFVHSQELLSRYPDLDAKGRERAIAKDLGAVFLVGIGGKLSDGHRHDVRAPDYDDWSTPSELGHAGLNGDILVWNPVLEDAFELSSMGIRVDADTLKHQLALTGDEDRLELEWHQALLRGEMPQTIGGGIGQSRLTMLLLQLPHIGQVQAGVWPAAVRESVPSLL",
        b"MA", b"ACDEFGHIKLMNPQRSTVWYXBZUOacd*-", b"G" * 17]


def hf_model(cfg: spec.ProstT5Config, w: dict) -> T5EncoderModel:
    hc = T5Config(vocab_size=cfg.n_vocab, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.n_layer,
                  num_heads=cfg.n_head, relative_attention_num_buckets=cfg.n_buckets,
                  relative_attention_max_distance=cfg.max_distance, dropout_rate=0.0, layer_norm_epsilon=cfg.eps,
                  feed_forward_proj="gated-gelu" if cfg.gated else "relu", is_encoder_decoder=False, use_cache=False)
    m = T5EncoderModel(hc).eval()
    sd = {}
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32).copy())
    sd["shared.weight"] = t(w["token_embd.weight"])
    sd["encoder.embed_tokens.weight"] = t(w["token_embd.weight"])
    for i in range(cfg.n_layer):
        p, q = f"enc.blk.{i}.", f"encoder.block.{i}.layer."
        sd[q + "0.layer_norm.weight"] = t(w[p + "attn_norm.weight"])
        sd[q + "0.SelfAttention.q.weight"] = t(w[p + "attn_q.weight"])
        sd[q + "0.SelfAttention.k.weight"] = t(w[p + "attn_k.weight"])
        sd[q + "0.SelfAttention.v.weight"] = t(w[p + "attn_v.weight"])
        sd[q + "0.SelfAttention.o.weight"] = t(w[p + "attn_o.weight"])
        if i == 0:
            sd[q + "0.SelfAttention.relative_attention_bias.weight"] = t(w[p + "attn_rel_b.weight"])
        sd[q + "1.layer_norm.weight"] = t(w[p + "ffn_norm.weight"])
        if cfg.gated:
            sd[q + "1.DenseReluDense.wi_0.weight"] = t(w[p + "ffn_gate.weight"])
            sd[q + "1.DenseReluDense.wi_1.weight"] = t(w[p + "ffn_up.weight"])
        else:
            sd[q + "1.DenseReluDense.wi.weight"] = t(w[p + "ffn_up.weight"])
        sd[q + "1.DenseReluDense.wo.weight"] = t(w[p + "ffn_down.weight"])
    sd["encoder.final_layer_norm.weight"] = t(w["enc.output_norm.weight"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in k or "shared" in k for k in missing), missing
    return m


def run(cfg, tag, out, seed=SEED, seqs=None, row_step=1):
    seqs = SEQS if seqs is None else seqs
    w = synth.make_weights(cfg, seed)
    m = hf_model(cfg, w)
    toks = spec.vocab_tokens(cfg.n_vocab)
    conv0 = torch.nn.Conv1d(cfg.d_model, cfg.cnn_hidden, cfg.cnn_kernel, padding=cfg.cnn_kernel // 2)
    conv1 = torch.nn.Conv1d(cfg.cnn_hidden, cfg.cnn_classes, cfg.cnn_kernel, padding=cfg.cnn_kernel // 2)
    with torch.no_grad():
        conv0.weight.copy_(torch.from_numpy(w["cnn.conv0.weight"].astype(np.float32)))
        conv0.bias.copy_(torch.from_numpy(w["cnn.conv0.bias"].astype(np.float32)))
        conv1.weight.copy_(torch.from_numpy(w["cnn.conv1.weight"].astype(np.float32)))
        conv1.bias.copy_(torch.from_numpy(w["cnn.conv1.bias"].astype(np.float32)))
    for n, seq in enumerate(seqs):
        ids = torch.from_numpy(spec.tokenize(seq, toks).astype(np.int64))[None]
        with torch.no_grad():
            hid = m(input_ids=ids).last_hidden_state[0]  # [T, d]
            x = hid[1:].T[None]  # prefix row dropped, </s> row kept (Rostlab predict_3Di_encoderOnly.py)
            logits = conv1(torch.relu(conv0(x)))[0].T[: len(seq)]
        out[f"{tag}_ids_{n}"] = ids[0].numpy().astype(np.int32)
        out[f"{tag}_hidden_{n}"] = hid.numpy()[::row_step]
        out[f"{tag}_logits_{n}"] = logits.numpy()


FULL_SEED, FULL_ROW_STEP = 1, 8


def full_sequences():
    """config 2's first sequence (350 aa) and a 1,200-aa sequence of the same composition (seed 22)."""
    aa, off = spec.synthetic_proteome("config2", n=1)
    rng = np.random.default_rng(22)
    letters = np.frombuffer(spec.AA_LETTERS.encode(), np.uint8)
    long = letters[rng.choice(20, size=1200, p=spec.AA_FREQ / spec.AA_FREQ.sum())].tobytes()
    return [aa[:int(off[1])].tobytes(), long]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    torch.manual_seed(0)
    here = os.path.dirname(os.path.abspath(__file__))
    if what in ("tiny", "all"):
        torch.set_num_threads(1)
        out = {"seqs": np.array([s.decode() for s in SEQS])}
        run(spec.TINY, "relu", out)
        run(spec.ProstT5Config(**{**spec.TINY.to_dict(), "gated": True}), "gated", out)
        path = os.path.join(here, "hf_t5_tiny.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")
    if what in ("full", "all"):
        torch.set_num_threads(os.cpu_count())
        seqs = full_sequences()
        out = {"seqs": np.array([s.decode() for s in seqs]), "row_step": np.int64(FULL_ROW_STEP), "seed": np.int64(FULL_SEED)}
        run(spec.FULL, "full", out, seed=FULL_SEED, seqs=seqs, row_step=FULL_ROW_STEP)
        path = os.path.join(here, "hf_t5_full.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
