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


# ---- tensor-core image (csrc/convnet_tc.cuh) -------------------------------------------------------
TC_CHUNK_ROWS, TC_ROW_BYTES = 192, 128


def _swizzle_rows(rows_u16):
    """rows_u16: (R, 64) uint16 (fp16 bits), R multiple of 8 -> SWIZZLE_128B image: 16-byte chunk c of
    row r is stored at chunk position c ^ (r % 8)."""
    R = rows_u16.shape[0]
    chunks = rows_u16.reshape(R, 8, 8)
    out = np.empty_like(chunks)
    r = np.arange(R)
    for c in range(8):
        out[r, c ^ (r & 7)] = chunks[r, c]
    return out.reshape(R, 64)


def _to16(x, dtype):
    """float array -> (uint16 bit patterns, float32 values after rounding) in fp16 or bf16"""
    if dtype == "fp16":
        h = np.asarray(x, np.float32).astype(np.float16)
        return h.view(np.uint16), h.astype(np.float32)
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16), t.to(torch.float32).numpy()


SPLIT_LAYERS = (1, 2, 3)  # csrc/convnet_tc.cuh SPLIT_FIRST..SPLIT_LAST


def pack_tc(folded, dtype="fp16", precise=False):
    """16-bit tensor-core blob (dtype "fp16" or "bf16"): 21 chunks (layer l, ky) of 192 rows [kx=2 | kx=1 | kx=0] x 64 co, each row
    64 ci fp16 (K-major, pre-swizzled); L0 rows are [W_hi(16 ci) | W_lo(16 ci) | 0]; then the fp32 tail
    bias[7][64], w7[4][64], b7[4], fcw[4][64], fcb[4].
    precise=True (B2048_PREC_FP16X2_TC, fp16 only): every chunk of layers 1..3 is followed by its lo chunk
    fp16(W - fp16(W)) -> 30 chunks."""
    if precise and dtype != "fp16":
        raise ValueError("the precise tensor-core mode is fp16 hi+lo")
    chunks = []
    for l, w in enumerate(folded["conv_w"]):  # (64 co, cin, 3, 3) float64
        for ky in range(3):
            rows = np.zeros((TC_CHUNK_ROWS, 64), np.float32)
            rows_lo = None
            for blk, kx in enumerate((2, 1, 0)):  # dx = +1, 0, -1  ->  output column x = s-1, s, s+1
                wk = w[:, :, ky, kx]  # (co, ci)
                if l == 0:
                    _, hi = _to16(wk, dtype)
                    _, lo = _to16(wk - hi.astype(np.float64), dtype)
                    rows[blk * 64:(blk + 1) * 64, 0:16] = hi
                    rows[blk * 64:(blk + 1) * 64, 16:32] = lo
                else:
                    rows[blk * 64:(blk + 1) * 64, :] = wk.astype(np.float32)
            bits, vals = _to16(rows, dtype)
            if not np.isfinite(vals).all():
                raise ValueError("folded weights overflow " + dtype)
            chunks.append(_swizzle_rows(bits).reshape(-1))
            if precise and l in SPLIT_LAYERS:
                lo_rows = np.zeros((TC_CHUNK_ROWS, 64), np.float64)
                for blk, kx in enumerate((2, 1, 0)):
                    lo_rows[blk * 64:(blk + 1) * 64, :] = w[:, :, ky, kx] - vals[blk * 64:(blk + 1) * 64, :].astype(np.float64)
                lo_bits, _ = _to16(lo_rows.astype(np.float32), dtype)
                chunks.append(_swizzle_rows(lo_bits).reshape(-1))
    wbytes = np.concatenate(chunks).view(np.uint8)
    tail = np.concatenate([np.stack(folded["conv_b"]).reshape(-1), folded["head_w"].reshape(-1), folded["head_b"],
                           folded["fc_w"].reshape(-1), folded["fc_b"]]).astype(np.float32)
    return torch.from_numpy(np.concatenate([wbytes, tail.view(np.uint8)]).copy())


def pack_blob(folded, precision):
    """Weight image for a precision name of network.PRECISIONS."""
    if precision == "fp32":
        return pack_fp32(folded)
    if precision == "fp16x2":
        return pack_tc(folded, "fp16", precise=True)
    if precision in ("fp16", "bf16"):
        return pack_tc(folded, precision)
    raise ValueError("unknown precision " + repr(precision))
