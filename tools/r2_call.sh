set -u
for tag in base STG SC W4 W8 ALL; do
  lib=$PWD/tools/micro/lib_$tag.so; [ $tag = base ] && lib=$PWD/probabilit_b200/libprobabilit_b200.so
  echo "$tag: $(PBL_LIB=$lib timeout 300 python tools/stage_times.py 1e8 16 3 0 2>&1 | grep total_ms | tail -1)  single: $(PBL_LIB=$lib timeout 300 python tools/stage_times.py 1e8 4 3 1 2>&1 | grep total_ms | tail -1)"
done
