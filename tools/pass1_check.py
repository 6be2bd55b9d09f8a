import sys, time, torch
sys.path.insert(0, '/root/repo')
from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.api import Detector
from locomouse_cpp_b200.types import bb_de_params
spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, 5120, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
p = bb_de_params(cfg, side_h=spec.side_h)
det.bounding_box_tm_de(frames, p)
t = time.perf_counter()
for _ in range(3): out = det.bounding_box_tm_de(frames, p)
dt = (time.perf_counter() - t) / 3
print(f"pass1: {dt*1e3:.2f} ms for 5120 frames -> {5120/dt:.0f} frames/s, {5120*680000/dt/1e9:.0f} GB/s algorithmic", out[2][:3].tolist())
