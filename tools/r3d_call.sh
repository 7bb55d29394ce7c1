# round-3 session call D: the GPU tests the -x run of call C did not reach (it stopped at the test of the experimental in-place lifting kernel, since removed)
timeout 50 python -m pytest $(cat tools/r3d_remaining_gpu_tests.txt) -q 2>&1 | tail -8 | tee gpurun_out/r3d_gpu_tests.txt
