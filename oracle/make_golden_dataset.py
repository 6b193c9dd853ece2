"""Golden vectors for the record reader (game_dataset.py:6-48), produced by the UNMODIFIED reference.

    python oracle/make_golden_dataset.py        (build container only; needs /root/reference)

Writes two synthetic self-play records in the reference's npz format (selfplay.py:55-58), runs the
reference's GameDataset / ConvGameDataset / OneHotConvGameDataset on them and stores inputs + outputs in
tests/golden/dataset_golden.npz."""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("B2048_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    scratch = tempfile.mkdtemp(prefix="b2048_ds_")
    os.chdir(scratch)
    sys.path.insert(0, REF)
    import game_dataset as gd  # the reference, unmodified

    rs = np.random.RandomState(7)
    eg = np.load(os.path.join(OUT, "engine_golden.npz"))
    recs = []
    for i, (a, b) in enumerate(((0, 40), (40, 97))):
        boards = np.concatenate([eg["pf_final"][a:b], eg["mv_boards"][a:b]]).astype(np.uint64)
        results = rs.randint(0, 60, size=(boards.size, 4)).astype(np.float32)
        results[np.arange(boards.size), rs.randint(0, 4, boards.size)] += 5.0  # row max > 0
        np.savez(os.path.join(scratch, str(i).zfill(5)), boards=boards, results=results, score=1234 + i)
        recs.append((boards, results))
    g = {"rec0_boards": recs[0][0], "rec0_results": recs[0][1], "rec1_boards": recs[1][0], "rec1_results": recs[1][1]}
    for name, cls in (("plain", gd.GameDataset), ("conv", gd.ConvGameDataset), ("onehot", gd.OneHotConvGameDataset)):
        ds = cls(scratch + "/", 0, 2, device="cpu", soft=3.0)
        g[f"{name}_boards"] = ds.boards.numpy()
        g[f"{name}_results"] = ds.results.numpy()
    ds = gd.GameDataset(scratch + "/", 1, 2, device="cpu", soft=1.5)
    g["soft15_results"] = ds.results.numpy()
    np.savez_compressed(os.path.join(OUT, "dataset_golden.npz"), **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
