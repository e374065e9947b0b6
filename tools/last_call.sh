out=gpurun_out/r02_wg4; mkdir -p $out
(time timeout 120 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "wgrad_building or all_parameter or cfg1_full") > $out/pytest_wgrad.log 2>&1
tail -4 $out/pytest_wgrad.log
(NSOS_WGRAD_W8=1 timeout 120 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "wgrad_building") > $out/pytest_wgrad_w8.log 2>&1
tail -1 $out/pytest_wgrad_w8.log
timeout 120 python tools/time_wgrad.py > $out/time_wgrad.txt 2>&1; cat $out/time_wgrad.txt | tail -22
