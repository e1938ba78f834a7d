#!/bin/bash
# ncu --set full of the first / last convolution stream kernels (one launch each) + stall breakdown
mkdir -p gpurun_out /tmp/ncu
timeout -s KILL 600 ncu --set full --clock-control none --import-source on \
    -k regex:"${KREGEX:-conv3x3_last_k|conv3x3_first_k}" -s ${SKIP:-2} -c ${COUNT:-2} -f -o /tmp/ncu/s2_small \
    python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 0 > gpurun_out/s2_ncu_small.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/ncu/s2_small.ncu-rep > gpurun_out/s2_ncu_small_summary.txt 2>&1
python - <<'PY' > gpurun_out/s2_ncu_small_stalls.txt 2>&1
import csv, subprocess
raw = subprocess.run(["ncu", "-i", "/tmp/ncu/s2_small.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in ("issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "l1tex__data_bank_conflicts", "smsp__warps_eligible", "sm__inst_executed_pipe_", "l1tex__t_bytes", "lts__t_sector_hit_rate", "smsp__cycles_active.avg", "achieved_occupancy", "sm__warps_active")):
            try:
                if float(v.replace(",", "")) == 0: continue
            except ValueError:
                pass
            print(f"{h} = {v} {u}")
PY
cp /tmp/ncu/s2_small.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/s2_small.ncu-rep
