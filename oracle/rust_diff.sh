#!/usr/bin/env bash
# SURVEY.md Appendix D, executable: pin the oracle (and through it the CUDA path) against the REAL reference the day a
# Rust toolchain is available.  Nothing here runs in the build image of this repository (no cargo / rustc / network).
#
#   oracle/rust_diff.sh REF_SRC DB_DIR PREPROC_DIR READS_ARGS... -- [-s SEED] [-@ T]
#     REF_SRC      checkout of tprodanov/locityper v1.7.2 with WFA2/ cloned (build.rs:5,28-40)
#     DB_DIR       locus database (`locityper target` output), PREPROC_DIR = `locityper preproc` output
#     READS_ARGS   the input arguments of `locityper genotype` (-i reads1.fq reads2.fq ...)
#
# Steps
#   1. patch the reference: rust/gpu.rs -> src/solvers/gpu.rs, the accessors of rust/reference_additions.rs, the dump
#      hook (6) in analyze_locus; build with `cargo build --release --features cuda` (or without the feature after
#      deleting the #![cfg] line: the dump only needs FlatLocus).
#   2. run `locityper genotype --debug 2 -s SEED -@ T` with LCTP_DUMP_LCTI=$OUT/lcti: every locus leaves its
#      solve::Data as a .lcti directory (tools/lcti.py) next to the reference's own sol.csv.br / sol_ext.csv.br.
#   3. for every locus: `python oracle/lcti_solve.py` re-runs the oracle from the dumped RNG state with the same T and
#      scheme and writes sol.csv / sol_ext.csv in the reference's row formats.
#   4. diff (rows sorted: their order inside a stage depends on thread timing in the reference).
# A clean diff pins, in one go: rand 0.10 bounded sampling / shuffle / Floyd order, xoshiro seeding and jumps, the
# candidate sort tie order, statrs ln_gamma (through the depth table compare below) and the t-test pruning.
set -euo pipefail
REF_SRC=$1; DB=$2; PREPROC=$3; shift 3
READS=(); while [[ $# -gt 0 && $1 != "--" ]]; do READS+=("$1"); shift; done; shift || true
SEED=12345; T=8
while [[ $# -gt 0 ]]; do case $1 in -s) SEED=$2; shift 2;; -@) T=$2; shift 2;; *) echo "unknown $1"; exit 2;; esac; done
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=${OUT:-$PWD/rust_diff_out}; mkdir -p "$OUT"

command -v cargo >/dev/null || { echo "cargo not found: this script needs a Rust toolchain (see header)"; exit 3; }
command -v brotli >/dev/null || { echo "brotli CLI needed to read *.csv.br"; exit 3; }

cp "$HERE/rust/gpu.rs" "$REF_SRC/src/solvers/gpu.rs"
echo ">> apply rust/reference_additions.rs blocks (1)-(6) to $REF_SRC by hand (six small insertions), then press enter"; read -r
(cd "$REF_SRC" && cargo build --release)

LCTP_DUMP_LCTI="$OUT/lcti" "$REF_SRC/target/release/locityper" genotype "${READS[@]}" -d "$DB" -p "$PREPROC" \
    -o "$OUT/ref" --debug 2 -s "$SEED" -@ "$T"

status=0
for d in "$OUT"/lcti/*/; do
    locus=$(basename "$d")
    python "$HERE/oracle/lcti_solve.py" "$d" --threads "$T" --out "$OUT/oracle/$locus"
    for f in sol sol_ext; do
        brotli -dc "$OUT/ref/loci/$locus/$f.csv.br" | sort > "$OUT/oracle/$locus/$f.ref.sorted"
        sort "$OUT/oracle/$locus/$f.csv" > "$OUT/oracle/$locus/$f.oracle.sorted"
        if diff -q "$OUT/oracle/$locus/$f.ref.sorted" "$OUT/oracle/$locus/$f.oracle.sorted" >/dev/null; then
            echo "[$locus] $f.csv identical"
        else
            echo "[$locus] $f.csv DIFFERS:"; diff "$OUT/oracle/$locus/$f.ref.sorted" "$OUT/oracle/$locus/$f.oracle.sorted" | head -20; status=1
        fi
    done
    # final call + probabilities
    python - "$OUT/ref/loci/$locus/res.json.gz" "$OUT/oracle/$locus/res.json" <<'PY' || status=1
import gzip, json, sys
a = json.load(gzip.open(sys.argv[1])); b = json.load(open(sys.argv[2]))
ok = a["genotype"] == b["genotype"] and [o["genotype"] for o in a["options"]] == [o["genotype"] for o in b["options"]]
print("res.json call/ranking identical:", ok); sys.exit(0 if ok else 1)
PY
done
exit $status
