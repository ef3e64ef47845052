#!/bin/bash
# pin_from_go.sh — pin the oracle (and the CUDA path) to the REAL reference, wherever a Go toolchain exists.
#
#   scripts/pin_from_go.sh /path/to/raytracer-go [--bridge]
#
# 1. copies raytracer_go_b200/go/parity_dump_test.go into a scratch copy of the reference's internal/ package
#    (the reference checkout itself is not modified);
# 2. runs it there: the reference's own Sphere.Hit, Quad.Hit, Aabb.Hit, World.Hit, BVH.Hit, Scatter/Emit, GetTexture,
#    ToGamma2/ToRGB, NewCamera, GetRay, reflect/refract/reflectance and Ray.GetColor evaluate tests/golden/pin_inputs.json
#    and write tests/golden/from_go/pin_outputs.json;
# 3. runs the pytest that requires the oracle (and, with a GPU, the device) to equal that file bit for bit.
# --bridge additionally compile-checks the cgo bridge (render_b200.go) against include/rt_b200.h and librt_b200.so.
# Needs Go >= 1.21 and the module cache (or network) for the reference's two dependencies; the image this repository is
# built in has neither, which is why the fixture is not committed yet.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:?usage: pin_from_go.sh /path/to/raytracer-go [--bridge]}"
command -v go >/dev/null || { echo "no Go toolchain on PATH" >&2; exit 2; }
WORK="$(mktemp -d)"; trap 'rm -rf "$WORK"' EXIT
cp -r "$REF"/. "$WORK"/
cp "$REPO/raytracer_go_b200/go/parity_dump_test.go" "$WORK/internal/"
mkdir -p "$REPO/tests/golden/from_go"
( cd "$WORK" && RT_B200_PIN_INPUTS="$REPO/tests/golden/pin_inputs.json" \
    RT_B200_PIN_OUTPUTS="$REPO/tests/golden/from_go/pin_outputs.json" go test ./internal -run TestB200ParityDump -count=1 -v )
if [ "${2:-}" = "--bridge" ]; then
  mkdir -p "$WORK/b200/include" "$WORK/b200/lib"
  cp "$REPO/include/rt_b200.h" "$WORK/b200/include/"
  cp "$REPO/raytracer_go_b200/csrc/librt_b200.so" "$WORK/b200/lib/"
  cp "$REPO/raytracer_go_b200/go/render_b200.go" "$WORK/internal/"
  ( cd "$WORK" && CGO_ENABLED=1 go vet ./internal && CGO_ENABLED=1 go build ./... ) && echo "cgo bridge compiles"
fi
cd "$REPO" && python -m pytest tests/test_pin_from_go.py -q -rs
