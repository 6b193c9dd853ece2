"""Forward-only sweep (config 5): tcgen05 vs fp32 kernels, accuracy on all golden boards + throughput."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
for tag in ("shipped", "random"):
    sd = {k[len(f"sd_{tag}/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith(f"sd_{tag}/")}
    m = pkg.network.ConvNet(64, 3); m.load_state_dict(sd); m.cuda().eval()
    boards = torch.from_numpy(ng["lg_boards"].view(np.int64).copy()).cuda()
    ref = ng[f"lg_{tag}"]
    for prec in ("fp32", "fp16"):
        lp = m.forward_boards(boards, precision=prec).cpu().numpy()
        d = np.abs(lp - ref)
        print(f"{tag} {prec}: max|d|={d.max():.5f} p99.9={np.quantile(d, 0.999):.5f} argmax={100*(lp.argmax(1)==ref.argmax(1)).mean():.3f}% "
              f"within1e-2={100*(d.max(1)<1e-2).mean():.3f}%", flush=True)
rep = boards.repeat(120)[: 1 << 20]
for prec, sizes in (("fp16", [1 << 10, 1 << 14, 1 << 17, 1 << 20]), ("fp32", [1 << 10, 1 << 14, 1 << 16])):
    for n in sizes:
        b = rep[:n].contiguous()
        m.forward_boards(b, precision=prec); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.forward_boards(b, precision=prec); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{prec} n={n}: {ms:.3f} ms  {n/ms*1e3:.3e} boards/s  {n/ms*1e3*5.128704e6/1e12:.1f} TFLOP/s(valid taps)", flush=True)
