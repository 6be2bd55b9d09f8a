import sys, time, torch
sys.path.insert(0, '/root/repo')
from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.api import Detector
spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, 2560, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
for streams in (2, 1, 2):
    det.set_option("streams", streams)
    for _ in range(2): r = det.detect_batch(frames, bx, bs, bb)
    t = time.perf_counter()
    for _ in range(3): r = det.detect_batch(frames, bx, bs, bb)
    dt = (time.perf_counter() - t) / 3
    tm, nl = det.last_timing()
    print(f"streams={streams}: wall {dt*1e3:.2f} ms -> {2560/dt:.0f} frames/s, device total {tm['total']:.2f} ms, checksum {r.checksum()}")
