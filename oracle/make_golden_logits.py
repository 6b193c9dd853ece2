"""Generate tests/golden/c64b3_logits_100k.npz: >= 1e5 on-distribution boards with the log-probs of the
UNMODIFIED reference network (network.py:78-87) for the shipped and the random-init c64b3 weights.

Run in the build container only (needs /root/reference, read-only):

    python oracle/make_golden_logits.py

Boards (SURVEY.md 4.3 / VERDICT r01 item 1) come from the reference's own loops under its own MT19937
randomness, seeded here:
  * NN-policy trajectories of the shipped net   (eval_nn.py:17-47 loop, every board the model is asked about)
  * NN-policy trajectories of the random-init net (same loop)
  * fixed-order L,U,D,R trajectories             (board.py:192-225 play_fixed order, BoardArray.move_batch)
  * the first-spawn boards of mcts_nn lines from mid-game origins (mcts_nn.py:19-22)
The state dicts are the ones already committed in c64b3_golden.npz (checked equal to the checkpoint / the
torch.manual_seed(0) initialisation of make_golden.py), so the fixture carries boards and log-probs only.

  boards      u64[N]      deduplicated, shuffled
  source      u8[N]       0 shipped-policy, 1 random-policy, 2 fixed-order, 3 mcts_nn line starts
  lg_shipped  f32[N,4]    reference log-probs, shipped weights
  lg_random   f32[N,4]    reference log-probs, random-init weights
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("B2048_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    golden = np.load(os.path.join(OUT, "c64b3_golden.npz"))
    scratch = tempfile.mkdtemp(prefix="b2048_ref_")
    os.chdir(scratch)
    sys.path.insert(0, REF)
    import random

    import board as rb  # the reference, unmodified
    import torch
    from network import ConvNet

    torch.set_num_threads(os.cpu_count() or 1)
    random.seed(20260101)
    np.random.seed(20260101)

    def model_of(tag):
        m = ConvNet(channels=64, blocks=3)
        pfx = f"sd_{tag}/"
        m.load_state_dict({k[len(pfx):]: torch.from_numpy(np.array(golden[k])) for k in golden.files if k.startswith(pfx)})
        return m.eval()

    shipped, rnd = model_of("shipped"), model_of("random")
    ckpt = os.path.join(REF, "models/20200213/0_3400_soft3.0c64b3_p10_bs2048lr0.1d0.0_s0_best.pt")
    ref_sd = torch.load(ckpt, map_location="cpu")
    for k, v in shipped.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k  # the fixture copy IS the checkpoint

    def encode(bs):  # the reference's inline encoder, mcts_nn.py:25-35 / eval_nn.py:28-40
        b = np.array(bs, dtype=np.uint64)
        data = []
        for _ in range(16):
            data.append(b & 0xF)
            b >>= 4
        t = torch.tensor(np.array(data), dtype=torch.float32).transpose(0, 1)
        t = [t == i for i in range(16)]
        return torch.stack(t, dim=1).float().view(-1, 16, 4, 4)

    def policy_trajectories(model, games, cap):
        """eval_nn.py:17-47 with every queried board recorded."""
        arr = rb.BoardArray([rb.generate_init_tiles() for _ in range(games)])
        seen = []
        with torch.no_grad():
            while arr.boards and len(seen) < cap:
                seen.extend(arr.boards)
                preds = torch.argsort(model(encode(arr.boards)), dim=1, descending=True)
                arr.move_batch(preds)
        return seen

    def fixed_trajectories(games):
        arr = rb.BoardArray([rb.generate_init_tiles() for _ in range(games)])
        seen = []
        while arr.boards:
            seen.extend(arr.boards)
            arr.move_batch([(0, 1, 3, 2)] * len(arr.boards))
        return seen

    parts = []
    a = policy_trajectories(shipped, 96, 90000)
    print("shipped-policy boards", len(a), flush=True)
    parts.append((0, a))
    b = policy_trajectories(rnd, 300, 30000)
    print("random-policy boards", len(b), flush=True)
    parts.append((1, b))
    c = fixed_trajectories(120)
    print("fixed-order boards", len(c), flush=True)
    parts.append((2, c))
    # first boards of mcts_nn lines (root move + spawn, mcts_nn.py:19-22) from mid-game origins
    d = []
    origins = [a[k] for k in np.random.choice(len(a), 400, replace=False)]
    for o in origins:
        for i in range(4):
            nb, _, m = rb.move(o, i)
            if m:
                d.extend(rb.generate_tile(nb) for _ in range(5))
    print("mcts_nn line starts", len(d), flush=True)
    parts.append((3, d))

    boards = np.concatenate([np.array(p, np.uint64) for _, p in parts])
    source = np.concatenate([np.full(len(p), s, np.uint8) for s, p in parts])
    boards, first = np.unique(boards, return_index=True)
    source = source[first]
    perm = np.random.permutation(boards.size)
    boards, source = boards[perm], source[perm]
    out = {"boards": boards, "source": source}
    with torch.no_grad():
        for tag, model in (("shipped", shipped), ("random", rnd)):
            lg = [model(encode(boards[i:i + 8192])).numpy() for i in range(0, boards.size, 8192)]
            out[f"lg_{tag}"] = np.concatenate(lg).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "c64b3_logits_100k.npz"), **out)
    print("c64b3_logits_100k.npz:", boards.size, "boards;", {int(s): int((source == s).sum()) for s in range(4)})


if __name__ == "__main__":
    main()
