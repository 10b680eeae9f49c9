for lib in libbathgpu.so libbathgpu_jointe.so; do echo "== $lib"; BATHGPU_FWD=3 BATHGPU_LIB=$PWD/bath_b200/$lib python scripts/j_sweep.py 2>&1 | grep "bump=0" | sed 's/fwd+bck.*//'; done
BATHGPU_LIB=$PWD/bath_b200/libbathgpu_jointe.so timeout 600 python -m pytest tests/test_gpu_fs_forward.py -x -q 2>&1 | tail -2
