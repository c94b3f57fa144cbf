set -u
bash tools/gpu_profile.sh r2n
ls -la gpurun_out | tail; du -sh gpurun_out
