#!/bin/bash
# One GPU-box visit: parity tests, bench line, and (optionally) the ncu launch list of one step plus an ncu --set full
# capture of the step's main kernels.
# usage (inside gpurun): bash tools/gpu_check.sh TAG [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ "$2" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "harp_timed/" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-e2e --no-cpu-baseline --no-standin \
    > gpurun_out/${TAG}_launches.log 2>&1
  timeout 700 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "harp_timed/" \
    -k regex:"tile_raster_kernel|shade_deferred_kernel|empty_tiles_kernel|face_setup_kernel|bin_fill_kernel|geo_fwd_kernel|geo_bwd_kernel|light_depth" \
    -c 12 -f -o gpurun_out/${TAG}_pixel \
    python bench.py --steps 1 --warmup 3 --eager --no-e2e --no-cpu-baseline --no-standin > gpurun_out/${TAG}_ncu.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.log
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e(u8)", round(d["e2e"]["value"]), "f32host", round(d["e2e_float32_host"]["value"]),
      "dom", r["kernel"][:20], round(r["kernel_ms"], 3), {k[:14]: round(v, 3) for k, v in r["other_kernel_ms"].items()}, d["counters"])
print("standin", d.get("gpu_eager_standin"), "dropin", d.get("dropin_level1"))
PY
