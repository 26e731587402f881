#!/bin/bash
# Turn the round-2 ncu captures (gpurun_out/r02_*.ncu-rep, tools/ncu_round2.sh) into the summaries committed under profiles/.
cd "$(dirname "$0")/.."
COMMIT=$(git rev-parse --short HEAD)
for r in 4k_tc batch64_tc batch64_butterfly 8k_tc; do
  [ -f gpurun_out/r02_$r.ncu-rep ] || continue
  python tools/summarize_ncu.py gpurun_out/r02_$r.ncu-rep profiles/r02_$r "tools/ncu_round2.sh ($r), code at commit $COMMIT" > /dev/null
done
cp gpurun_out/r02_launches_uhd4k.csv profiles/r02_launches_uhd4k.csv 2>/dev/null
# executed-instruction histograms per opcode (from the captures) and static SASS evidence (from the shipped library)
{
  echo "# Executed warp instructions per opcode, batch of 64 x 1080p (69120 strips), from gpurun_out/r02_batch64_tc.ncu-rep"
  for k in k_fused k_strip k_merge; do
    ncu -i gpurun_out/r02_batch64_tc.ncu-rep --page source --csv --kernel-name regex:$k --print-source sass 2>/dev/null > /tmp/_sass_$k.csv
    echo; echo "## $k (per strip)"; python tools/sass_hist.py /tmp/_sass_$k.csv 69120 2>/dev/null | head -28
  done
  echo; echo "## k_fused_blocks<false> (butterfly transform, dct_mode 2), same workload"
  ncu -i gpurun_out/r02_batch64_butterfly.ncu-rep --page source --csv --kernel-name regex:k_fused --print-source sass 2>/dev/null > /tmp/_sass_bf.csv
  python tools/sass_hist.py /tmp/_sass_bf.csv 69120 2>/dev/null | head -24
} > profiles/r02_sass_executed_hist.txt
{
  echo "# Static SASS of jpeg_image_compression_b200/libjpegb200.so (cuobjdump -sass), commit $COMMIT"
  echo "# mnemonic counts per kernel: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG / UBLKCP, packed fp32 -> FFMA2/FADD2/FMUL2"
  cuobjdump -sass jpeg_image_compression_b200/libjpegb200.so | awk '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UBLKCP|FFMA2|FADD2|FMUL2|HADD2|IDP|ATOMS|RED|VABSDIFF4|SYNCS|UTCATOMSWS)/) { split($i,a,"."); c[fn" "a[1]]++ } }
    END { for (k in c) print k, c[k] }' | sort | grep -E "k_fused|k_strip|k_merge"
} > profiles/r02_sass_static.txt
python - <<'PY'
import json, subprocess, csv, io
def dram(rep, kern):
    raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    def tob(v, unit): return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    for r in rows[2:]:
        if kern in r[h.index('Kernel Name')]:
            i, j = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
            return tob(r[i], u[i]) + tob(r[j], u[j])
commit = subprocess.run("git rev-parse --short HEAD", shell=True, capture_output=True, text=True).stdout.strip()
out = {"note": "DRAM bytes (read + write) of k_fused_blocks per launch from ncu --set full captures; bench.py reports them as roofline.traffic with traffic_source = static_ncu"}
a = dram("gpurun_out/r02_4k_tc.ncu-rep", "k_fused")
b = dram("gpurun_out/r02_batch64_tc.ncu-rep", "k_fused")
if a: out["uhd4k"] = {"dram_bytes_per_launch": int(a), "capture": "profiles/r02_4k_tc_ncu_full_summary.md", "commit": commit}
if b: out["batch1080p"] = {"dram_bytes_per_launch": int(b / 64), "per": "image (capture: 64 images per launch)", "capture": "profiles/r02_batch64_tc_ncu_full_summary.md", "commit": commit}
json.dump(out, open("profiles/k1_dram_traffic.json", "w"), indent=1)
print(out)
PY
ls profiles | grep r02
