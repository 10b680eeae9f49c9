set -x
mkdir -p gpurun_out
python -m pytest tests/test_calibration.py -m gpu -x -q 2>&1 | tail -8
python scripts/calibrate_time.py > gpurun_out/calibrate_time.json 2> gpurun_out/calibrate_time.err; tail -3 gpurun_out/calibrate_time.err; cat gpurun_out/calibrate_time.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
