#!/bin/bash
# End-of-round check on one GPU: smoke(), the driver's bench command (timed by wall clock), a 200-step bench line, the reference arm.
TAG=${1:-r2q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
T0=$SECONDS
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err
echo "driver-style bench exit $? wall $((SECONDS-T0)) s"; python tools/show_extras.py gpurun_out/${TAG}_bench20.json
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; python tools/show_extras.py gpurun_out/${TAG}_bench.json
T0=$SECONDS
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
echo "reference arm exit $? wall $((SECONDS-T0)) s"; tail -c 400 gpurun_out/${TAG}_reference_arm.json
if [ "${2:-}" = "tests" ]; then
  timeout 1400 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
