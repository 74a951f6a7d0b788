#!/bin/bash
# Build step of the CPU reference arm: copies the UNMODIFIED reference modules that make up the hot path into
# oracle/_ref/ (git-ignored like a built .so; it travels to the GPU box with the gpurun snapshot, where
# /root/reference does not exist).  Nothing is edited; `bench.py --impl reference` imports these files as they are
# and patches `.cuda()` to identity at run time, exactly as tests/golden/make_golden.py does.
#   training_code/utils_my.py            group_points_3DV*, global_contrast, circle_contrast
#   training_code/cn3d_model_conbag.py   PointNet_Plus_fine / PointNet_Plus
#   training_code/cn3D_data_set.py       farthest_point_sampling_fast (its source lines are exec'd; the module needs imageio)
# usage: tools/vendor_reference.sh [reference root, default /root/reference]
set -eu
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
DST="$HERE/oracle/_ref"
if [ ! -d "$REF/training_code" ]; then
    echo "vendor_reference: $REF/training_code not found (GPU box?): keeping whatever is in $DST" >&2
    exit 0
fi
mkdir -p "$DST"
for f in utils_my.py cn3d_model_conbag.py cn3D_data_set.py; do
    cp "$REF/training_code/$f" "$DST/$f"
done
( cd "$DST" && sha256sum utils_my.py cn3d_model_conbag.py cn3D_data_set.py > SHA256SUMS )
echo "vendor_reference: $(ls "$DST" | tr '\n' ' ')-> $DST"
