"""Precise-mode root evaluations: single launches after idle vs back-to-back launches on one / two streams (sustained clocks)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200"); eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16x2")
m.load_state_dict({k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")})
m.cuda().eval()
org = eng.init_boards(256, 4242, 0)
for prec in ("fp16x2", "fp16"):
    blob = m.blob(prec)
    eng.mcts_nn_lines(blob, prec, org, 50, 400, 0); torch.cuda.synchronize()
    for nstreams in (1, 2):
        for reps in (1, 3, 8):
            time.sleep(0.5)
            cur = torch.cuda.current_stream()
            lanes = [torch.cuda.Stream() for _ in range(nstreams)] if nstreams > 1 else [cur]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in lanes: st.wait_stream(cur)
            rs = []
            for k in range(reps):
                with torch.cuda.stream(lanes[k % len(lanes)]):
                    rs.append(eng.mcts_nn_lines(blob, prec, org, 50, 500 + k, k * 256))
            for st in lanes: cur.wait_stream(st)
            e1.record(); torch.cuda.synchronize()
            mv = sum(int(r["total_moves"].item()) for r in rs)
            print(f"{prec}: {reps} evaluation(s) on {nstreams} stream(s): {e0.elapsed_time(e1) / reps:.1f} ms each, {mv / e0.elapsed_time(e1) * 1e3:.4g} moves/s", flush=True)
