"""Turn the ncu reports / launch list in gpurun_out/ into the committed text summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def raw(rep):
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers']
R1 = ("prof_ctc_r1b", "prof_ctcpair_r1", "prof_gemmtc_r1", "prof_sweepcl_r1", "prof_sweep_r1")
R2 = ("prof_c2_r2", "prof_ctc_r2", "prof_c3_r2", "prof_sweeptc_nocoop_r2")
reps = [r for r in (R2 if tag == "r2" else R1) if os.path.exists(os.path.join(G, r + ".ncu-rep"))]
traffic = {}
with open(os.path.join(P, "ncu_%s_summary.txt" % tag), "w") as f:
    if tag == "r2":
        f.write("ncu --set full --clock-control none --import-source on captures on B200 (sm_100a); commands in tools/profile_r2.sh:\n"
                "  prof_c2_r2             one C2 step (H=512, B=32, T=200): sweep_cluster_kernel, ctc_pair_kernel, every gemm_tc_kernel\n"
                "  prof_ctc_r2            ctc_warp_kernel alone, B=8192 x C1 shape (814 MB > L2)\n"
                "  prof_c3_r2             gemm_tc_kernel launches of a C3 step (H=1024, B=128, T=800)\n"
                "  prof_sweeptc_nocoop_r2 sweep_tc_kernel of a C3 step (CTCB_SWEEP_TC_COOP=0: ncu cannot replay the cooperative+cluster launch)\n\n")
    else:
      f.write("ncu --set full --clock-control none --import-source on captures on B200 (sm_100a), one launch each:\n"
            "  -k regex:ctc_warp -s 1 -c 1       python tools/prof_ctc.py                                   (B=8192 x C1 shape, 814 MB > L2)\n"
            "  -k regex:ctc_pair -s 1 -c 1       python bench.py --steps 2 --warmup 1 --no-cpu-baseline      (inside the C2 step, B=32)\n"
            "  -k regex:gemm_tc_kernel -s 6 -c 2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline      (inside the C2 step)\n"
            "  -k regex:sweep_cluster -s 2 -c 1  python bench.py --steps 2 --warmup 1 --no-cpu-baseline      (inside the C2 step)\n"
            "  -k regex:sweep_kernel  -s 2 -c 1  (general counter-barrier kernel, captured before the cluster kernel became the default)\n\n")
    for rep in reps:
        rows = raw(rep)
        if len(rows) < 3:
            continue
        hdr, unit = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            f.write("== %s   [%s]\n" % (name[:100], rep))
            for w in WANT:
                if w in hdr:
                    f.write("   %-72s %-16s %s\n" % (w, unit[hdr.index(w)], r[hdr.index(w)]))
            def num(k):
                return float(r[hdr.index(k)]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[unit[hdr.index(k)]]
            if r[hdr.index('gpu__time_duration.sum')] in ('0', '-nan', 'nan'):
                continue
            key = name.split('(')[0].replace('void ', '')
            traffic.setdefault(key, num('dram__bytes_read.sum') + num('dram__bytes_write.sum'))
            f.write("\n")
    f.write("SASS evidence (cuobjdump -sass stanford-ctc_b200/libctcb200.so):\n"
            "  sweep_tc_kernel          UTCHMMA (tcgen05.mma, SS and TS forms), UTMALDG.3D/.4D (TMA), LDTM / STTM (tcgen05.ld / .st), UTCBAR, LDS via mapa (DSMEM)\n"
            "  gemm_tc_kernel           UTCHMMA (tcgen05.mma), UTMALDG.2D (TMA), LDTM.x32 (tcgen05.ld), UTCBAR (tcgen05.commit)\n"
            "  sweep_cluster_kernel     STAS (st.async into cluster shared memory), SYNCS.PHASECHK.TRANS64.TRYWAIT (mbarrier), FFMA2\n"
            "  ctc_warp/ctc_pair_kernel LDGSTS (cp.async), REDUX / CREDUX, DADD / DMUL, ATOMS.ADD\n")
json.dump(traffic, open(os.path.join(P, "ncu_traffic_%s.json" % tag), "w"), indent=1)

rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    k = r[ik].split('(')[0].replace('void ', '')[:52]; agg[k][0] += 1; agg[k][1] += float(r[iv].replace(',', ''))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, "launches_%s_summary.txt" % tag), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none %s python bench.py %s\n"
            "(cold-cache and serialised: compare SHARES, not absolutes)\n\n" % (("-s 150 -c 400", "--headline-only --no-cpu-baseline --steps 3 --warmup 3   [CTCB_NO_GRAPH=1]") if tag == "r2" else ("-c 400", "--steps 2 --warmup 1 --no-cpu-baseline")))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-54s n=%4d total=%9.1f us avg=%8.1f us share=%.3f\n" % (k, c, t / 1e3, t / c / 1e3, t / tot))
subprocess.run(["cp", os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "launches_%s.csv" % tag)])
print(open(os.path.join(P, "launches_%s_summary.txt" % tag)).read()[:1400]); print(traffic)
