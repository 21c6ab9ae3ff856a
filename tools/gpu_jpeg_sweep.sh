python -m pytest tests/test_jpeg_gpu.py -x -q -m gpu 2>&1 | tail -3
export CV_B200_LIB=chess_vision_b200/libchessvision_b200_exp.so
for cfg in "CV_JPEG_CHUNK=0" "CV_JPEG_MCUS=8" "CV_JPEG_MCUS=4" "CV_JPEG_MCUS=16" "CV_JPEG_MCUS=4 CV_JPEG_CHUNK=128" "CV_JPEG_MCUS=2 CV_JPEG_CHUNK=64 CV_JPEG_ROUNDS=10"; do
  echo "== $cfg"
  env $cfg CV_JPEG_TRACE=1 python tools/gpu_jpeg_bench.py 2>&1 | grep -v "entropy_on_host=True\|(host)\|H2D tables" | tail -12 | grep "chunked\|entropy decode\|n=" 
done
