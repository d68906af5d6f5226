"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (round tag as argv[1])."""
import csv, io, subprocess, json, collections, shutil, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
cmd = "python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline"
rows = [r for r in csv.reader(open(f'gpurun_out/{tag}_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0]; t = float(r[-1])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if 'smb::' in k)
lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ({cmd})",
         "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
         "# share_of_library = share among this library's kernels (the rest is torch: L2 flush fills, int8 GEMM peak probe, RNG)",
         "kernel,launches,total_ns,share_of_all,share_of_library"]
for k, v in agg.items():
    lines.append(f"{k[:110]},{v[0]},{v[1]:.0f},{v[1]/tot:.4f},{(v[1]/ours if 'smb::' in k else 0):.4f}")
open(f'profiles/{tag}_launches_summary.csv', 'w').write("\n".join(lines) + "\n"); print("\n".join(lines))
shutil.copy(f'gpurun_out/{tag}_launches.csv', f'profiles/{tag}_launches.csv')
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'launch__registers_per_thread',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum',
        'smsp__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'pcie__write_bytes.sum', 'pcie__read_bytes.sum',
        'lts__t_sectors_aperture_sysmem_op_write.sum', 'lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_st.sum']
mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}
for kern in ('score', 'decide', 'runner_up'):
    rep = f'gpurun_out/{tag}_prof_{kern}.ncu-rep'
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    out = {'_source': f'ncu --set full --clock-control none --import-source on -k regex:{kern} -s 4 -c 1 {cmd}',
           '_kernel': rr[2][4] if len(rr) > 2 else ''}
    for h, u, v in zip(rr[0], rr[1], rr[2]):
        for w in want:
            if h.endswith(w) and w not in out:
                out[w] = {'unit': u, 'value': v}
    json.dump(out, open(f'profiles/{tag}_{kern}_kernel_ncu_full.json', 'w'), indent=1)
    print(kern, {k: v for k, v in out.items() if not k.startswith('_')})
    if kern == 'score':
        tr = float(out['dram__bytes_read.sum']['value']) * mul[out['dram__bytes_read.sum']['unit']] + \
             float(out['dram__bytes_write.sum']['value']) * mul[out['dram__bytes_write.sum']['unit']]
        json.dump({'dram_bytes_per_launch': tr, 'source': f'profiles/{tag}_score_kernel_ncu_full.json (ncu --set full, '
                   'score_tcgen05_kernel, 855 pairs of 8192x8192)'}, open('profiles/roofline_traffic.json', 'w'))
        print('traffic', tr)
