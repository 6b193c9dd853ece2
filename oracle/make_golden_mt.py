"""Distribution samples from the UNMODIFIED reference under its OWN randomness (CPython MT19937
`randrange`, seed 12345 at import: board.py:8-12) -- the KS-test counterpart of the stream-exact goldens.

    python oracle/make_golden_mt.py            (build container only; needs /root/reference)

  fixed:  4000 x board.play_fixed()                      -> final board, score, move count
  nn:     eval_nn.eval_nn(shipped c64b3, number=300)     -> scores, moves (eval_nn.py:4-58, via its npz output)
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("B2048_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    scratch = tempfile.mkdtemp(prefix="b2048_mt_")
    os.chdir(scratch)
    os.makedirs("models", exist_ok=True)
    sys.path.insert(0, REF)
    import board as rb
    import eval_nn as re_
    import torch
    from network import ConvNet

    g = {}
    res = np.array([rb.play_fixed() for _ in range(4000)], dtype=np.uint64)
    g["fixed_final"], g["fixed_score"], g["fixed_moves"] = res[:, 0], res[:, 1], res[:, 2]
    m = ConvNet(channels=64, blocks=3)
    m.load_state_dict(torch.load(os.path.join(REF, "models/20200213/0_3400_soft3.0c64b3_p10_bs2048lr0.1d0.0_s0_best.pt"),
                                 map_location="cpu"))
    re_.eval_nn(m, name="mt_sample", number=300, device="cpu")
    z = np.load("models/mt_sample.npz")
    g["nn_score"], g["nn_moves"] = z["scores"].astype(np.int64), np.asarray(z["moves"], np.int64)
    np.savez_compressed(os.path.join(OUT, "reference_mt_samples.npz"), **g)
    print("fixed mean log", np.log10(g["fixed_score"].astype(float) + 1).mean(), "moves", g["fixed_moves"].mean(),
          "| nn mean log", np.log10(g["nn_score"] + 1.0).mean(), "moves", g["nn_moves"].mean())


if __name__ == "__main__":
    main()
