"""Stall-reason samples and scheduler statistics of the first kernel in an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
st = {}
for j, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        st[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = int(float(rows[2][j].replace(',', '')))
    if h in ('smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active'):
        print(h, rows[2][j])
tot = sum(st.values())
print('stall samples:', ', '.join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
