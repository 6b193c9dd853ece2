"""SM clock / power while the NN-policy playout kernel (or the forward) runs back to back for a few seconds."""
import importlib, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3, precision="fp16"); m.load_state_dict(sd); m.cuda().eval()
blob = m.blob("fp16")
samples, stop = [], False
def sample():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.2)
for name, fn in (("mcts_nn 256x4x50", lambda: pkg.engine.mcts_nn_lines(blob, "fp16", pkg.engine.init_boards(256, 3, 0), 50, 9, 0)),
                 ("forward 1M boards", lambda: m.forward_boards(pkg.engine.init_boards(1 << 20, 5, 0)))):
    samples.clear(); stop = False
    th = threading.Thread(target=sample); th.start()
    t0 = time.time(); n = 0
    while time.time() - t0 < 4.0:
        fn(); n += 1
        if n % 8 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize(); stop = True; th.join()
    print(name, "launches", n, "samples (MHz, W, reasons):", samples[3:-1][:12], flush=True)
