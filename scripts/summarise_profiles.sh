#!/bin/bash
# gpurun_out/*.ncu-rep -> profiles/ncu_search_r02.txt (summary counters + per-source-line stall samples per capture)
out=profiles/ncu_search_r02.txt
: > $out
for pair in "prof_r02_raw10m|beam_search_kernel, raw rows, 10M x 128 fp32 L2, 10k queries at the bench operating point (python bench.py, one timed step)" \
            "prof_r02_ef280|beam_search_kernel, raw rows, hard variant (L=32) 1M x 128, ef 280 (its 0.95 point), BEFORE double hashing / destination-tile merge" \
            "prof_r02_ef280_final|beam_search_kernel, the same launch AFTER the round-2 changes (double hashing, destination-tile merge, six rows in flight)" \
            "prof_r02_c5_ef144|beam_search_kernel, one C5 shard: 12.5M x 96 fp32 L2, ef 144" \
            "prof_r02_sq8_c3|beam_search_kernel, SQ8 codes, C3 1M x 960 L2, ef 180" \
            "prof_r02_sq4_c4|beam_search_kernel, SQ4 codes, C4 shape 2M x 768 inner product, top-100, ef 160" \
            "prof_r02b_ef280|beam_search_kernel, hard variant at ef 280 with the two-choice bucket table (final code of round 2)" \
            "prof_r02b_sq8_c3|beam_search_kernel, SQ8 C3 1M x 960 ef 180, final code (bucket table, VABSDIFF4 decode, bulk L2 prefetch of later passes)" \
            "prof_r02b_sq4_c4|beam_search_kernel, SQ4 2M x 768 inner product top-100 ef 160, final code (bucket table)"; do
  f=${pair%%|*}; title=${pair#*|}
  [ -f gpurun_out/$f.ncu-rep ] || continue
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > /tmp/_raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
  echo "==== $f: $title" >> $out
  echo "(ncu --set full --clock-control none --import-source on, one launch; columns of the first block: metric, unit, value)" >> $out
  python scripts/ncu_lines.py /tmp/_src.csv /tmp/_raw.csv 2>&1 | head -34 >> $out
  python - >> $out <<'PY'
import csv
rows = list(csv.reader(open('/tmp/_raw.csv')))
h, u, r = rows[0], rows[1], rows[-1]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.avg.per_cycle_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__block_size', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
for i, n in enumerate(h):
    if n in want:
        print(f"{n} [{u[i]}] = {r[i]}")
PY
  echo >> $out
done
# DRAM bytes of the headline launch for bench.py's roofline.traffic (only used when the run's n / ef match)
python - <<'PY'
import csv, json, subprocess, os
rep = "gpurun_out/prof_r02_raw10m.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, r = rows[0], rows[1], rows[-1]
    def val(name):
        i = h.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[i]]
        return int(float(r[i]) * scale)
    plain = json.load(open("gpurun_out/r02_prof_bench_plain.json"))
    out = {"main": {"n": 10_000_000, "dim": 128, "nq": 10_000, "ef": plain["config"]["ef"], "dram_bytes_read": val("dram__bytes_read.sum"),
                    "dram_bytes_write": val("dram__bytes_write.sum"), "source": "ncu --set full, one timed launch of python bench.py (profiles/ncu_search_r02.txt)"}}
    json.dump(out, open("profiles/traffic_r02.json", "w"), indent=1)
    print(out)
PY
wc -l $out
