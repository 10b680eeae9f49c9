for lib in libbathgpu.so libbathgpu_jointb.so; do echo "== $lib"; BATHGPU_LIB=$PWD/bath_b200/$lib python scripts/j_sweep.py 2>&1 | grep "bump=0" | sed 's/fwd  *[0-9.]*  //'; done
BATHGPU_LIB=$PWD/bath_b200/libbathgpu_jointb.so timeout 600 python -m pytest tests/test_gpu_fs_backward.py -x -q 2>&1 | tail -2
