python -m pytest tests -m gpu -x -q -k "mc" 2>&1 | tail -15 > gpurun_out/s3b_tests.log
for v in libv3d libv3d_d384 libv3d_d512x1 libv3d_d1024x1; do echo $v; V3D_LIB_PATH=video-3d-llm_b200/lib/$v.so python tools/time_mc.py 10; done > gpurun_out/s3b_time.log 2>&1
echo hash-only >> gpurun_out/s3b_time.log; V3D_MC_DENSE=0 python tools/time_mc.py 10 >> gpurun_out/s3b_time.log 2>&1
cat gpurun_out/s3b_tests.log gpurun_out/s3b_time.log
