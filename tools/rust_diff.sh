#!/bin/bash
# pins the oracle against the real reference on the first box with cargo.
# The script itself lives next to the oracle it pins (oracle/rust_diff.sh: builds the reference with the `.lcti` dump
# patch of rust/reference_additions.rs, runs `locityper genotype -s SEED -@ T --debug 2`, re-runs the oracle on the
# dumped solve::Data with oracle/lcti_solve.py and diffs sol.csv / sol_ext.csv / depth.csv row for row); this is the
# entry point at the path the verdict names.
exec "$(dirname "$0")/../oracle/rust_diff.sh" "$@"
