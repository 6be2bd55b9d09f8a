"""Device timeline of the two-stream pipeline: start/end of every stage of the last sub-batch in each slot."""
import sys

import torch

sys.path.insert(0, '/root/repo')
from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402

spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
N = 2048
frames, bx, bs, bb = synth.make_video(spec, N, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
names = ["start", "minmax_end", "prep_end", "corr_end", "tail_end", "nms_end", "pair_end", "d2h_end"]
for streams in (2, 1):
    det.set_option("streams", streams)
    for _ in range(3):
        det.detect_batch(frames, bx, bs, bb)
    print(f"streams={streams} total {det.last_timing()[0]['total']:.3f} ms (4 sub-batches of 512; slot 0 ran #2, slot 1 ran #3)")
    for slot in (0, 1):
        print(f"  slot {slot}: " + "  ".join(f"{n}={det.info(f'stage_t_{slot}_{k}'):.3f}" for k, n in enumerate(names)))
