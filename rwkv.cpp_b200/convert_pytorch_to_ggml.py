"""PyTorch RWKV checkpoint (.pth) -> rwkv.cpp model file, byte-compatible with the reference converter
(reference python/convert_pytorch_to_ggml.py:28-159; container format docs/FILE_FORMAT.md:10-41).

    python rwkv.cpp_b200/convert_pytorch_to_ggml.py model.pth model-FP16.bin FP16

The result is what `rwkv_init_from_file` of this engine (and of the reference) loads; quantise it afterwards with
`rwkv_quantize_model_file`. The per-architecture parameter rewrites are table-driven here (RULES below), one row per
reference line, so the parity test can name the rule a differing tensor went through:

  v4    time_decay := -exp(raw)                                                    (:123-124)
  v5.1  time_decay := exp(-exp(raw)) as (H,1,1);  time_first := exp(raw) as (H,1,1) (:110-116)
  v5.2  time_decay := exp(-exp(raw)) with a trailing axis; time_faaaa gets a trailing axis (:111-121)
  v6    time_faaaa trailing axis; time_maa_w1 / time_decay_w1 / time_decay_w2 transposed; time_maa_w2 transposed(1,2);
        time_decay reshaped (H,-1,1)                                                (:100-108)
  v7    the six att.x_* vectors of a layer concatenated (in checkpoint order) into att.x_rwkvag; LoRA pairs w1/w2, a1/a2,
        v1/v2, g1/g2 transposed                                                     (:51-66, 91-98)
  all   names containing `.time_` are squeezed first (:88-89); FP16 output keeps FP32 for 1-D tensors and for names
        containing .time_ .k_k .k_a .r_k .x_rwkvag .x_k .w0 .a0 .v0                (:126-135)
Tensors are written in state-dict order, dimensions reversed (ggml order), no padding.
"""
import argparse
import struct
from typing import Callable, Dict, List, Tuple

import numpy as np

MAGIC = 0x67676D66
FILE_VERSION = 101
KEEP_FP32_MARKERS = (".time_", ".k_k", ".k_a", ".r_k", ".x_rwkvag", ".x_k", ".w0", ".a0", ".v0")
V7_LORA_MARKERS = (".w1", ".w2", ".a1", ".a2", ".v1", ".v2", ".g1", ".g2")


def detect_architecture(state_dict) -> Tuple[int, int]:
    """(major, minor) from the parameter names, same precedence as the reference (:36-39, 41-50)."""
    if "blocks.0.att.k_k" in state_dict:
        return 7, 0
    if "blocks.0.att.time_maa_x" in state_dict:
        return 6, 0
    if "blocks.0.att.ln_x.weight" in state_dict:
        return (5, 2) if "blocks.0.att.gate.weight" in state_dict else (5, 1)
    return 4, 0


def count_layers(state_dict) -> int:
    n = 0
    while f"blocks.{n}.ln1.weight" in state_dict:
        n += 1
    if n == 0:
        raise ValueError("not an RWKV checkpoint: blocks.0.ln1.weight is missing")
    return n


def _merge_v7_token_shift(state_dict) -> Dict:
    """att.x_r, x_w, x_k, x_v, x_a, x_g of each layer -> one att.x_rwkvag, concatenated along dim 0 in checkpoint order and
    placed where the layer's first x_ tensor stood."""
    import torch
    merged: Dict = {}
    for name, tensor in state_dict.items():
        if "att.x_" in name:
            layer = int(name.split(".")[1])
            key = f"blocks.{layer}.att.x_rwkvag"
            merged[key] = torch.cat([merged[key], tensor], dim=0) if key in merged else tensor
        else:
            merged[name] = tensor
    return merged


def _rules(arch: Tuple[int, int], n_head: int) -> List[Tuple[Callable[[str], bool], Callable]]:
    """(predicate on the tensor name, rewrite) pairs, applied in order after the `.time_` squeeze."""
    import torch
    major, minor = arch
    if major == 7:
        return [(lambda k: any(s in k for s in V7_LORA_MARKERS), lambda t: t.transpose(0, 1))]
    if major == 6:
        return [
            (lambda k: ".time_faaaa" in k, lambda t: t.unsqueeze(-1)),
            (lambda k: ".time_maa_w1" in k or ".time_decay_w" in k, lambda t: t.transpose(0, 1)),
            (lambda k: ".time_maa_w2" in k, lambda t: t.transpose(1, 2)),
            (lambda k: ".time_decay" in k and "_w" not in k, lambda t: t.reshape(n_head, -1, 1)),
        ]
    if major == 5:
        decay = (lambda t: torch.exp(-torch.exp(t)).unsqueeze(-1)) if minor >= 2 else (lambda t: torch.exp(-torch.exp(t)).reshape(-1, 1, 1))
        return [
            (lambda k: ".time_decay" in k, decay),
            (lambda k: ".time_first" in k, lambda t: torch.exp(t).reshape(-1, 1, 1)),
            (lambda k: ".time_faaaa" in k, lambda t: t.unsqueeze(-1)),
        ]
    return [(lambda k: ".time_decay" in k, lambda t: -torch.exp(t))]


def transform_tensor(name: str, tensor, arch: Tuple[int, int], n_head: int, want_fp16: bool):
    """One checkpoint tensor -> the tensor the file stores (fp32 or fp16 torch tensor, PyTorch dimension order)."""
    t = tensor.float()
    if ".time_" in name:
        t = t.squeeze()
    for applies, rewrite in _rules(arch, n_head):
        if applies(name):
            t = rewrite(t)
    if want_fp16 and t.dim() > 1 and not any(s in name for s in KEEP_FP32_MARKERS):
        t = t.half()
    return t


def write_state_dict(state_dict, dest_path: str, data_type: str, verbose: bool = False) -> None:
    """Same entry point name and arguments as the reference (:28)."""
    import torch
    want_fp16 = data_type in ("FP16", "float16")
    if not want_fp16 and data_type not in ("FP32", "float32"):
        raise ValueError(f"data_type must be FP16 or FP32, got {data_type}")
    emb = state_dict["emb.weight"]
    n_vocab, n_embed = int(emb.shape[0]), int(emb.shape[1])
    n_layer = count_layers(state_dict)
    arch = detect_architecture(state_dict)
    if verbose:
        print("Detected RWKV v%d%s" % (arch[0], ".%d" % arch[1] if arch[0] in (5, 6, 7) else ""))
    if arch[0] == 7:
        state_dict = _merge_v7_token_shift(state_dict)
    n_head = int(state_dict["blocks.0.att.time_faaaa"].shape[0]) if arch[0] == 6 else 0
    with open(dest_path, "wb") as out:
        out.write(struct.pack("<6i", MAGIC, FILE_VERSION, n_vocab, n_embed, n_layer, 1 if want_fp16 else 0))
        for name, tensor in state_dict.items():
            t = transform_tensor(name, tensor, arch, n_head, want_fp16)
            key = name.encode("utf-8")
            dims = list(t.shape)
            out.write(struct.pack("<3i", len(dims), len(key), 1 if t.dtype == torch.float16 else 0))
            out.write(struct.pack("<%di" % len(dims), *reversed(dims)))     # ggml order: fastest dimension first
            out.write(key)
            np.ascontiguousarray(t.detach().contiguous().numpy()).tofile(out)
            if verbose:
                print(f"Writing {name}, shape {tuple(dims)}, type {t.dtype}")


def main() -> None:
    import torch
    ap = argparse.ArgumentParser(description="Convert an RWKV checkpoint in PyTorch format to an rwkv.cpp compatible file")
    ap.add_argument("src_path", help="Path to PyTorch checkpoint file")
    ap.add_argument("dest_path", help="Path to rwkv.cpp checkpoint file, will be overwritten")
    ap.add_argument("data_type", help="Data type, FP16 or FP32", choices=["FP16", "FP32", "float16", "float32"], default="FP16", nargs="?")
    args = ap.parse_args()
    print(f"Reading {args.src_path}")
    state_dict = torch.load(args.src_path, map_location="cpu")
    write_state_dict(state_dict, args.dest_path, args.data_type, verbose=True)
    print("Done")


if __name__ == "__main__":
    main()
