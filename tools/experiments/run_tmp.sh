python -m pytest tests -m gpu -x -q -k "mc" 2>&1 | grep -v Warn | tail -2 > gpurun_out/s3j_tests.log
python tools/time_mc.py 10 > gpurun_out/s3j_time.log 2>&1
V3D_LIB_PATH=video-3d-llm_b200/lib/libv3d_d1024x1.so python tools/time_mc.py 10 >> gpurun_out/s3j_time.log 2>&1
cat gpurun_out/s3j_tests.log; grep -o "'lib[^,]*\|voxelize_ms_best[^,]*" gpurun_out/s3j_time.log
