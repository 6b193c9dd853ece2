"""Host side of the c64b3 policy network: BatchNorm folding and weight-blob packing.

The reference network is ConvNet(channels=64, blocks=3) (network.py:43-87).  For inference
(eval mode) every BatchNorm2d is folded into the preceding bias-free conv in float64:
    w' = w * gamma / sqrt(running_var + 1e-5),   b' = beta - running_mean * gamma / sqrt(...)
(raw running_var reaches 5e7 in the shipped checkpoint, so folding must precede any rounding:
SURVEY.md Appendix C).  The folded weights are packed into the device images the CUDA kernels
read ("blobs"); layouts are documented beside the kernels (csrc/convnet_fp32.cuh,
csrc/convnet_tc.cuh).
"""
import numpy as np
import torch

CONVS = [("in_block.0", "in_block.1")] + [(f"blocks.{b}.{c}", f"blocks.{b}.{c + 1}") for b in range(3) for c in (0, 3)]


def fold_state_dict(sd, channels=64, blocks=3):
    """state_dict (reference key names) -> folded float64 numpy arrays."""
    if channels != 64 or blocks != 3:
        raise NotImplementedError("the B200 kernels implement c64b3 = ConvNet(channels=64, blocks=3)")

    def get(k):
        return sd[k].detach().to("cpu", torch.float64).numpy()

    def fold(conv, bn):
        w = get(conv + ".weight")
        s = get(bn + ".weight") / np.sqrt(get(bn + ".running_var") + 1e-5)
        return w * s[:, None, None, None], get(bn + ".bias") - get(bn + ".running_mean") * s

    out = {"conv_w": [], "conv_b": []}
    for conv, bn in CONVS:
        w, b = fold(conv, bn)
        out["conv_w"].append(w)  # (64, Cin, 3, 3)
        out["conv_b"].append(b)
    w7, b7 = fold("out_block.0", "out_block.1")
    out["head_w"], out["head_b"] = w7.reshape(4, 64), b7
    out["fc_w"], out["fc_b"] = get("policy.weight"), get("policy.bias")  # (4, 64), (4,)
    if out["conv_w"][0].shape != (64, 16, 3, 3) or out["fc_w"].shape != (4, 64):
        raise ValueError("state_dict is not a c64b3 ConvNet")
    return out


def pack_fp32(folded):
    """fp32 blob of csrc/convnet_fp32.cuh: L0 [16 ci][9 tap][64 co] + bias, L1..6 [64 ci][9][64 co] + bias, head."""
    parts = []
    for w, b in zip(folded["conv_w"], folded["conv_b"]):
        cin = w.shape[1]
        parts.append(np.transpose(w.reshape(64, cin, 9), (1, 2, 0)).reshape(-1))  # [ci][tap][co]
        parts.append(b)
    parts += [folded["head_w"].reshape(-1), folded["head_b"], folded["fc_w"].reshape(-1), folded["fc_b"]]
    return torch.from_numpy(np.concatenate(parts).astype(np.float32))
