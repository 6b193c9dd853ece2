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
SPLIT_CHANNELS = 32       # csrc/convnet_tc.cuh SPLIT_NK * 16: input channels whose lo terms the precise mode adds


def onehot_np(boards):
    """uint64 boards -> float32 (N,16,4,4) one-hot planes (channel = cell exponent, cell k at (k // 4, k % 4): mcts_nn.py:25-35)."""
    b = np.ascontiguousarray(boards, np.uint64)
    nib = np.stack([(b >> np.uint64(4 * k)) & np.uint64(0xF) for k in range(16)], axis=1).astype(np.int64)  # (N,16 cells)
    x = np.zeros((b.size, 16, 16), np.float32)
    np.put_along_axis(x, nib[:, None, :], 1.0, axis=1)
    return x.reshape(-1, 16, 4, 4)


def split_channel_order(folded, calib_boards):
    """Channel orders for the precise tensor-core mode.  Its correction passes (x_lo * w_hi, x_hi * w_lo on layers 1..3) run over
    the first SPLIT_CHANNELS input channels only, so the channels whose rounding errors matter most -- largest
    E[x_c^2] * sum(w[:, c]^2) on the calibration boards -- must come first.  Returns (stream order, mid order): the residual
    stream (outputs of layers 0, 2, 4, 6 = inputs of layers 1, 3, 5 and of the head) shares one order, chosen for layers 1 and 3
    together; the output of layer 1 (input of layer 2) has its own.  tools/precision_study3.py: max |dlogp| 0.0067 on the
    150k-board fixture against 0.0052 with all 64 channels, at half the correction work (0.0078 in the shipped configuration,
    which also drops layer 3's activation correction)."""
    import torch.nn.functional as F

    x = torch.from_numpy(onehot_np(calib_boards))
    w = [torch.from_numpy(np.asarray(a, np.float32)) for a in folded["conv_w"]]
    bias = [torch.from_numpy(np.asarray(a, np.float32)).view(1, -1, 1, 1) for a in folded["conv_b"]]
    with torch.no_grad():
        x0 = F.relu(F.conv2d(x, w[0], padding=1) + bias[0])
        h1 = F.relu(F.conv2d(x0, w[1], padding=1) + bias[1])
        x2 = F.relu(F.conv2d(h1, w[2], padding=1) + bias[2] + x0)

        def importance(act, l):
            v = act.pow(2).mean((0, 2, 3)) * w[l].pow(2).sum((0, 2, 3))
            return v / v.sum().clamp_min(1e-30)

        stream = importance(x0, 1) + importance(x2, 3)
        mid = importance(h1, 2)
    return (np.argsort(-stream.numpy(), kind="stable"), np.argsort(-mid.numpy(), kind="stable"))


def permute_channels(folded, stream, mid):
    """The same network with its channels reordered: position j of the residual stream holds old channel stream[j], position j
    of layer 1's output holds old channel mid[j].  Sums run over the same terms in another order: results agree to fp32 rounding."""
    w, b = [np.array(a) for a in folded["conv_w"]], [np.array(a) for a in folded["conv_b"]]
    w[0], b[0] = w[0][stream], b[0][stream]
    w[1], b[1] = w[1][mid][:, stream], b[1][mid]
    w[2], b[2] = w[2][stream][:, mid], b[2][stream]
    for l in (3, 5):
        w[l] = w[l][:, stream]
    for l in (4, 6):
        w[l], b[l] = w[l][stream], b[l][stream]
    out = dict(folded)
    out["conv_w"], out["conv_b"] = w, b
    out["head_w"] = np.array(folded["head_w"])[:, stream]
    return out


def pack_tc(folded, dtype="fp16", precise=False, calib_boards=None):
    """16-bit tensor-core blob (dtype "fp16" or "bf16"): 21 chunks (layer l, ky) of 192 rows [kx=2 | kx=1 | kx=0] x 64 co, each row
    64 ci fp16 (K-major, pre-swizzled); L0 rows are [W_hi(16 ci) | W_lo(16 ci) | 0]; then the fp32 tail
    bias[7][64], w7[4][64], b7[4], fcw[4][64], fcb[4].
    precise=True (B2048_PREC_FP16X2_TC, fp16 only): layers 1..3 carry two lo chunks fp16(W - fp16(W)) of their first 32 input
    channels ([ky 0 | ky 1] after the second chunk, [ky 2 | 0] after the third) -> 27 chunks; the channels are reordered (split_channel_order on `calib_boards`) so that the kernel's
    correction passes over the first SPLIT_CHANNELS input channels cover the ones that matter."""
    if precise and dtype != "fp16":
        raise ValueError("the precise tensor-core mode is fp16 hi+lo")
    if precise:
        if calib_boards is None or len(calib_boards) < 256:
            raise ValueError("the precise tensor-core image needs calibration boards (network.calibration_boards)")
        folded = permute_channels(folded, *split_channel_order(folded, calib_boards))
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
                # lo weights fp16(W - fp16(W)) of the first SPLIT_CHANNELS input channels; a 64-column row holds the halves of
                # two taps side by side: chunk X = [ky 0 | ky 1] follows the layer's second chunk, Y = [ky 2 | 0] its third
                if ky != 1:
                    lo_rows = np.zeros((TC_CHUNK_ROWS, 64), np.float64)
                col = SPLIT_CHANNELS if ky == 1 else 0
                for blk, kx in enumerate((2, 1, 0)):
                    lo_rows[blk * 64:(blk + 1) * 64, col:col + SPLIT_CHANNELS] = (
                        w[:, :SPLIT_CHANNELS, ky, kx] - vals[blk * 64:(blk + 1) * 64, :SPLIT_CHANNELS].astype(np.float64))
                if ky >= 1:
                    lo_bits, _ = _to16(lo_rows.astype(np.float32), dtype)
                    chunks.append(_swizzle_rows(lo_bits).reshape(-1))
    wbytes = np.concatenate(chunks).view(np.uint8)
    tail = np.concatenate([np.stack(folded["conv_b"]).reshape(-1), folded["head_w"].reshape(-1), folded["head_b"],
                           folded["fc_w"].reshape(-1), folded["fc_b"]]).astype(np.float32)
    return torch.from_numpy(np.concatenate([wbytes, tail.view(np.uint8)]).copy())


def pack_blob(folded, precision, calib_boards=None):
    """Weight image for a precision name of network.PRECISIONS (calib_boards: uint64 boards, needed by "fp16x2")."""
    if precision == "fp32":
        return pack_fp32(folded)
    if precision == "fp16x2":
        return pack_tc(folded, "fp16", precise=True, calib_boards=calib_boards)
    if precision in ("fp16", "bf16"):
        return pack_tc(folded, precision)
    raise ValueError("unknown precision " + repr(precision))
