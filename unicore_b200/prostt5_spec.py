"""ProstT5 (T5-encoder + CNN 3Di head) model description shared by the loader, the synthetic-weight
writer and the tests: hyper-parameters, gguf tensor names, vocabulary and the synthetic proteome
generators of BASELINE.json's configs.

Sources: tensor names follow the ``t5encoder`` architecture of the gguf package
(gguf/constants.py MODEL_TENSORS[T5ENCODER]); the arithmetic they feed is specified in SURVEY.md §8a
p0-p11 (HF modeling_t5.py for the encoder, Rostlab ``predict_3Di_encoderOnly.py`` for the CNN head).
The names of the CNN-head tensors inside Foldseek's own ``prostt5-f16.gguf`` are unknown here
(SURVEY.md open question Q1); ``CNN_NAME_ALIASES`` lists the candidates the loader accepts.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np

ARCH = "t5encoder"
THREE_DI_ALPHABET = "ACDEFGHIKLMNPQRSTVWY"  # class index -> 3Di letter (p11)
WEIGHT_FILE = "prostt5-f16.gguf"  # [REF src/modules/createdb.rs:148]
OLD_WEIGHT_FILES = ("cnn.safetensors", "model/cnn.safetensors")  # [REF src/modules/createdb.rs:144]


@dataclass(frozen=True)
class ProstT5Config:
    n_layer: int = 24
    d_model: int = 1024
    n_head: int = 32
    d_kv: int = 128
    d_ff: int = 16384
    n_vocab: int = 150
    n_buckets: int = 32
    max_distance: int = 128
    eps: float = 1e-6
    gated: bool = False
    cnn_hidden: int = 32
    cnn_classes: int = 20
    cnn_kernel: int = 7

    @property
    def d_inner(self) -> int:
        return self.n_head * self.d_kv

    def flops_per_seq(self, L: int) -> float:
        """Algorithmic FLOPs of one sequence of L residues (T = L + 2 tokens)."""
        T = L + 2
        ffn = (3 if self.gated else 2) * self.d_model * self.d_ff
        per_tok = self.n_layer * 2 * (4 * self.d_model * self.d_inner + ffn)
        attn = self.n_layer * 4 * self.d_inner  # QK^T + PV, per (query, key) pair
        head = 2 * self.cnn_kernel * (self.d_model * self.cnn_hidden + self.cnn_hidden * self.cnn_classes)
        return T * (per_tok + attn * T) + head * L

    def to_dict(self):
        return asdict(self)


FULL = ProstT5Config()
# small enough to commit as a fixture, still exercising every code path (d_kv stays 128: the
# attention kernel is specialised on the ProstT5 head size)
TINY = ProstT5Config(n_layer=2, d_model=128, n_head=2, d_kv=128, d_ff=256)


def vocab_tokens(n_vocab: int = 150) -> list[str]:
    """ProstT5 sentencepiece vocabulary order ([EXT])."""
    toks = ["<pad>", "</s>", "<unk>"]
    toks += ["▁" + c for c in "ALGVSREDTIPKFQNYMHWCXBOUZ"]
    toks += [f"<extra_id_{99 - i}>" for i in range(100)]
    toks += ["▁" + c for c in "acdefghiklmnpqrstvwy"]
    toks += ["<fold2AA>", "<AA2fold>"]
    assert len(toks) == 150
    if n_vocab < len(toks):
        raise ValueError("vocabulary smaller than the ProstT5 token set")
    toks += [f"<unused_{i}>" for i in range(n_vocab - len(toks))]
    return toks


def tensor_shapes(cfg: ProstT5Config) -> list[tuple[str, tuple, str]]:
    """(name, numpy shape, dtype) of every tensor, in file order."""
    out = [("token_embd.weight", (cfg.n_vocab, cfg.d_model), "f2")]
    for i in range(cfg.n_layer):
        p = f"enc.blk.{i}."
        out.append((p + "attn_norm.weight", (cfg.d_model,), "f4"))
        out.append((p + "attn_q.weight", (cfg.d_inner, cfg.d_model), "f2"))
        out.append((p + "attn_k.weight", (cfg.d_inner, cfg.d_model), "f2"))
        out.append((p + "attn_v.weight", (cfg.d_inner, cfg.d_model), "f2"))
        out.append((p + "attn_o.weight", (cfg.d_model, cfg.d_inner), "f2"))
        if i == 0:
            out.append((p + "attn_rel_b.weight", (cfg.n_buckets, cfg.n_head), "f4"))
        out.append((p + "ffn_norm.weight", (cfg.d_model,), "f4"))
        if cfg.gated:
            out.append((p + "ffn_gate.weight", (cfg.d_ff, cfg.d_model), "f2"))
        out.append((p + "ffn_up.weight", (cfg.d_ff, cfg.d_model), "f2"))
        out.append((p + "ffn_down.weight", (cfg.d_model, cfg.d_ff), "f2"))
    out.append(("enc.output_norm.weight", (cfg.d_model,), "f4"))
    out.append(("cnn.conv0.weight", (cfg.cnn_hidden, cfg.d_model, cfg.cnn_kernel), "f2"))
    out.append(("cnn.conv0.bias", (cfg.cnn_hidden,), "f4"))
    out.append(("cnn.conv1.weight", (cfg.cnn_classes, cfg.cnn_hidden, cfg.cnn_kernel), "f2"))
    out.append(("cnn.conv1.bias", (cfg.cnn_classes,), "f4"))
    return out


# candidate names for the CNN head (first match wins); index 0 is what this repo writes
CNN_NAME_ALIASES = {
    "conv0.weight": ["cnn.conv0.weight", "cnn.0.weight", "classifier.0.weight", "conv0.weight"],
    "conv0.bias": ["cnn.conv0.bias", "cnn.0.bias", "classifier.0.bias", "conv0.bias"],
    "conv1.weight": ["cnn.conv1.weight", "cnn.3.weight", "classifier.3.weight", "conv1.weight"],
    "conv1.bias": ["cnn.conv1.bias", "cnn.3.bias", "classifier.3.bias", "conv1.bias"],
}


def metadata(cfg: ProstT5Config, name: str = "ProstT5-synthetic") -> dict:
    a = ARCH
    return {
        "general.architecture": a,
        "general.name": name,
        f"{a}.block_count": cfg.n_layer,
        f"{a}.embedding_length": cfg.d_model,
        f"{a}.feed_forward_length": cfg.d_ff,
        f"{a}.attention.head_count": cfg.n_head,
        f"{a}.attention.key_length": cfg.d_kv,
        f"{a}.attention.value_length": cfg.d_kv,
        f"{a}.attention.layer_norm_rms_epsilon": float(cfg.eps),
        f"{a}.attention.relative_buckets_count": cfg.n_buckets,
        f"{a}.vocab_size": cfg.n_vocab,
        "tokenizer.ggml.model": "t5",
        "tokenizer.ggml.tokens": vocab_tokens(cfg.n_vocab),
        "tokenizer.ggml.eos_token_id": 1,
        "tokenizer.ggml.padding_token_id": 0,
        "tokenizer.ggml.unknown_token_id": 2,
    }


def config_from_gguf(g) -> ProstT5Config:
    """Hyper-parameters from gguf metadata + tensor shapes (mirrors csrc/model.cu load_config)."""
    m = g.meta
    a = m.get("general.architecture", ARCH)
    d_model = int(m[f"{a}.embedding_length"])
    n_layer = int(m[f"{a}.block_count"])
    n_head = int(m[f"{a}.attention.head_count"])
    d_kv = int(m.get(f"{a}.attention.key_length", g.infos["enc.blk.0.attn_q.weight"].shape[0] // n_head))
    d_ff = int(m.get(f"{a}.feed_forward_length", g.infos["enc.blk.0.ffn_up.weight"].shape[0]))
    rel = g.infos["enc.blk.0.attn_rel_b.weight"].shape
    c0 = _first(g, CNN_NAME_ALIASES["conv0.weight"])
    c1 = _first(g, CNN_NAME_ALIASES["conv1.weight"])
    return ProstT5Config(
        n_layer=n_layer, d_model=d_model, n_head=n_head, d_kv=d_kv, d_ff=d_ff,
        n_vocab=g.infos["token_embd.weight"].shape[0], n_buckets=int(rel[0]),
        eps=float(m.get(f"{a}.attention.layer_norm_rms_epsilon", 1e-6)),
        gated="enc.blk.0.ffn_gate.weight" in g.infos,
        cnn_hidden=g.infos[c0].shape[0], cnn_classes=g.infos[c1].shape[0], cnn_kernel=g.infos[c0].shape[2])


def _first(g, names):
    for n in names:
        if n in g.infos:
            return n
    raise KeyError(f"none of {names} found in {g.path}")


def cnn_tensor(g, key: str) -> np.ndarray:
    return g.tensor(_first(g, CNN_NAME_ALIASES[key]))


# ---------------------------------------------------------------------------------------------
# tokenisation (p1)
# ---------------------------------------------------------------------------------------------
def residue_lut(tokens: list[str], map_rare_to_x: bool = True) -> np.ndarray:
    """256-entry byte -> token id table: upper-cased residue ``c`` -> id of "▁C"; anything without a
    token -> "▁X"; ProstT5's preprocessing maps U, Z, O, B to X as well."""
    idx = {t: i for i, t in enumerate(tokens)}
    x_id = idx["▁X"]
    lut = np.full(256, x_id, np.int32)
    for b in range(256):
        ch = chr(b).upper()
        if len(ch) != 1 or not ("A" <= ch <= "Z"):
            continue
        if map_rare_to_x and ch in "UZOB":
            continue
        lut[b] = idx.get("▁" + ch, x_id)
    return lut


def special_ids(tokens: list[str]) -> tuple[int, int]:
    idx = {t: i for i, t in enumerate(tokens)}
    return idx["<AA2fold>"], idx["</s>"]


def tokenize(seq: bytes, tokens: list[str]) -> np.ndarray:
    prefix, eos = special_ids(tokens)
    lut = residue_lut(tokens)
    ids = np.empty(len(seq) + 2, np.int32)
    ids[0] = prefix
    ids[1:-1] = lut[np.frombuffer(seq, np.uint8)]
    ids[-1] = eos
    return ids


# ---------------------------------------------------------------------------------------------
# synthetic proteomes
# ---------------------------------------------------------------------------------------------
AA_LETTERS = "LEAKISVGDRTNFPYQMHCW"
AA_FREQ = np.array([9.28, 8.07, 7.70, 7.30, 6.77, 6.69, 6.67, 6.47, 5.90, 5.67, 4.87, 4.44, 4.28, 3.79, 3.34, 3.32,
                    2.20, 1.56, 0.84, 0.84])
WORKLOAD_SEED = 20261017


def synthetic_lengths(workload: str, n: int | None = None) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(WORKLOAD_SEED))
    if workload == "config2":  # 256 x 350 aa
        return np.full(n or 256, 350, np.int64)
    if workload == "config4":  # 100k seqs, lognormal lengths clipped to 64..1024
        n = n or 100000
        return np.clip(np.round(rng.lognormal(np.log(260.0), 0.65, n)), 64, 1024).astype(np.int64)
    if workload == "config5":  # 4k seqs, 2000..4000 aa
        n = n or 4000
        return rng.integers(2000, 4001, n).astype(np.int64)
    raise ValueError(workload)


def synthetic_proteome(workload: str, n: int | None = None):
    """Returns (aa bytes [sum L] uint8, offsets [n+1] uint64) with residues i.i.d. from the
    example/data composition."""
    lens = synthetic_lengths(workload, n)
    offsets = np.zeros(len(lens) + 1, np.uint64)
    offsets[1:] = np.cumsum(lens)
    rng = np.random.Generator(np.random.PCG64(WORKLOAD_SEED + 1))
    letters = np.frombuffer(AA_LETTERS.encode(), np.uint8)
    aa = letters[rng.choice(len(letters), size=int(offsets[-1]), p=AA_FREQ / AA_FREQ.sum())]
    return np.ascontiguousarray(aa), offsets
