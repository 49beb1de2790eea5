# compute-sanitizer over smoke() and two small golden cases (incl. a banded host call and a local-ensemble pass):
#   gpurun --timeout 1500 -- 'bash profiles/sanitize.sh'   -> gpurun_out/sanitize_{memcheck,racecheck,synccheck}.log
mkdir -p gpurun_out
cat > /tmp/san_run.py <<'PY'
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as g
import stif_b200
from stif_b200 import synthetic as synth
g.smoke()
w = synth.make_weights(2, True)
for mode in ("bf16", "fp32"):
    dec = stif_b200.STIFQueryDecoder(0, mode=mode); dec.load_weights(w)
    lat, fr = synth.make_inputs(2, 2, 12, 10, 1.0)                                  # odd_b2_stress geometry
    a = dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [torch.tensor([[0.25], [0.75]])], (37, 53))
    lat, fr = synth.make_inputs(1, 1, 16, 16, 0.05)
    e = dec.decode_localensemble(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [0.25], (104, 104))   # x6.5
    lat, fr = synth.smooth_inputs(3, 1, 40, 48, 0.05)
    band = dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [0.5], None, rows=(40, 88), halo=16)
    if mode == "bf16":
        dec.host_pipeline(bands=4, halo=8)
        h = dec.decode_host(lat, fr, [0.0, 0.5, 0.75], (160, 192))                    # banded host pipeline, 3 timesteps
        u = dec.decode_host(lat, fr, [0.5], (160, 192), uint8=True)
    torch.cuda.synchronize()
    print(mode, "ok", float(a.abs().mean()), float(e.abs().mean()))
    dec.close()
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_$tool.log | tail -5
done
