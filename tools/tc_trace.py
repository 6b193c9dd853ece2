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
print("per set-layer (cycles): issue window = mma:issued_dy2 - mma:act_ready (tensor floor 3840 for a 64-channel layer, 1920 for L0);"
      " drain = wrk:acc_ready - mma:issued_dy2; epilogue = wrk:epi_done - wrk:acc_ready")
for l in range(7):
    for s_ in range(2):
        r = t[l * 2 + s_]
        print(f"L{l} set{s_}: issue window {int(r[1]-r[0]):6d}  drain {int(r[2]-r[1]):6d}  epilogue {int(r[3]-r[2]):6d}")
r0, r1 = t[0], t[13]
print("forward of both sets, first act_ready -> last epilogue done:", int(max(t[12][3], t[13][3]) - min(t[0][0], t[1][0])), "cycles")
