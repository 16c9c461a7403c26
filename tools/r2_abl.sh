set -x
python tests/ablate.py "D L,dgrad 3x3,vgg" 2>&1 | tee gpurun_out/r2_abl_halo.log
