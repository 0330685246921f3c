#!/bin/sh
# Applies the overlay to a checkout of blefaudeux/rusty-marcher:   integration/apply.sh /path/to/rusty-marcher
# Copies build.rs, src/ffi.rs, src/flat.rs into engine/, vendors this repository's include/ and csrc/ as engine/rm_b200/
# (what build.rs compiles) and patches the eight files that change.  `cargo run --release` then renders on the B200.
set -e
REF=${1:?usage: apply.sh /path/to/rusty-marcher}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
cp "$HERE/engine/build.rs" "$REF/engine/build.rs"
cp "$HERE/engine/src/ffi.rs" "$HERE/engine/src/flat.rs" "$REF/engine/src/"
mkdir -p "$REF/engine/rm_b200/rusty_marcher_b200"
cp -r "$ROOT/include" "$REF/engine/rm_b200/"
cp -r "$ROOT/rusty_marcher_b200/csrc" "$REF/engine/rm_b200/rusty_marcher_b200/"
for p in "$HERE"/engine/patches/*.patch; do
    patch -d "$REF" -p0 < "$p"
done
echo "overlay applied to $REF/engine"
