"""GPU: patch-embedding forward (+ the rest of the stage: addend, TDL finish) fused vs three-kernel path, CUDA-event
timing of the model's eval-free training forward being too coarse: times dcv_embed_fwd through the module's own plan
by running the training forward of a depth-0 ... simply: the full-channel bench breakdown tags of both settings."""
import json
import os
import subprocess
import sys

for f in (sys.argv[1:] or ("0", "1", "2")):
    env = dict(os.environ, DCV_EMBED_FUSED=f)
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-eager", "--steps", "10", "--warmup", "3"], env=env,
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    kb = d["kernel_breakdown_ms_per_step"]
    print(f"fused={f}: value {d['value']:.0f} img/s, full-channel {d['full_channels']['value']:.0f} img/s, embed tags (ms/step, full channels):",
          {k: kb.get(k) for k in ("im2col", "embed_gemm", "tdl", "embed_misc")}, flush=True)
