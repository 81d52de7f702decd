#!/bin/bash
# Builds liboswald_cuda.so of another commit into build/exp/ (for A/B runs on one GPU box; build/ travels
# with gpurun, git ignores it).   usage: bash tools/build_at.sh <commit> [name]   ->  build/exp/liboswald_<name>.so
set -e
COMMIT=${1:?commit}; NAME=${2:-$(git rev-parse --short "$COMMIT")}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" worktree add -q "$TMP/src" "$COMMIT"
make -C "$TMP/src" -j8 lib > "$TMP/build.log" 2>&1 || { tail -20 "$TMP/build.log"; exit 1; }
mkdir -p "$ROOT/build/exp"
cp "$TMP/src/oswald_b200/liboswald_cuda.so" "$ROOT/build/exp/liboswald_$NAME.so"
git -C "$ROOT" worktree remove --force "$TMP/src"
rm -rf "$TMP"
echo "$ROOT/build/exp/liboswald_$NAME.so"
