# ncu --set full of the detector's non-GEMM kernels (one launch each) from a short bench run; CSV pages -> gpurun_out/
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'roi_align|head_post|rpn_topk|nms_mask|nms_scan|maxpool|stem_canvas' --launch-skip 40 -c 14 -o /tmp/nongemm python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_nongemm.log 2>&1
ncu -i /tmp/nongemm.ncu-rep --page raw --csv > gpurun_out/nongemm_raw.csv 2>/dev/null
ls -la /tmp/nongemm.ncu-rep
