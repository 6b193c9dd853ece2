"""Timeline of one forward of CTA 0 (needs a build with -DB2048_TRACE=<iteration>)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3); m.load_state_dict(sd); m.cuda().eval()
boards = torch.from_numpy(ng["lg_boards"].view(np.int64).copy()).cuda().repeat(16)[: 4736 * 24].contiguous()
for _ in range(2):
    m.forward_boards(boards, precision="fp16")
torch.cuda.synchronize()
lib = C.CDLL(pkg.LIB_PATH)
buf = (C.c_longlong * 512)()
lib.b2048_debug_trace(buf, 512)
t = np.array(buf[:], dtype=np.int64).reshape(-1, 8)[:14]
t0 = t[t > 0].min()
names = ["mma:act_ready", "mma:issued_dy2", "wrk:acc_ready", "wrk:epi_done", "w_full dy0", "w_full dy1", "w_full dy2", "-"]
for l in range(7):
    for s in range(2):
        r = t[l * 2 + s]
        print(f"L{l} set{s}: " + "  ".join(f"{names[i]}={int(r[i]-t0) if r[i] else -1:7d}" for i in range(7 if s == 0 else 4)))
tt = np.array(buf[:], dtype=np.int64)[256:320].reshape(2, 32)[:, :16]
for s in range(2):
    d = np.diff(tt[s])
    print(f"set{s} L3 dy1 per-MMA issue intervals (slab order 0,3,1,2; N=128,128,192,192):", d.tolist(), "first-last", int(tt[s][-1] - tt[s][0]))
e = np.array(buf[:], dtype=np.int64)[384:384 + 56].reshape(14, 4)
for l in range(1, 6):
    for s_ in range(2):
        r = t[l * 2 + s_]; x = e[l * 2 + s_]
        print(f"L{l} set{s_}: acc_ready->chunks_done {int(x[0]-r[2])}  tc_fence_before {int(x[1]-x[0])}  fence.proxy.async {int(x[2]-x[1])}  arrive {int(r[3]-x[2])}")
