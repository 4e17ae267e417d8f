"""Turns gpurun_out/<launch csv> and <.ncu-rep> into the markdown summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches_summary.md "title" "command"
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_summary.md "title" "command"
"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_sectors.sum.per_second', 'lts__t_sectors.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_requests_srcunit_tex_op_red.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def launches(src, dst, title, cmd):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:70]
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f"# {title}\n\nCommand: `{cmd}`\n\n(cold-cache, serialised launches: compare SHARES with bench.py's CUDA-event "
                "shares, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
            f.write(f"| `{k}` | {c} | {t:.1f} | {t / tot * 100:.2f}% |\n")
        f.write(f"\nTotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches.\n")


def full(src, dst, title, cmd):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f"# {title}\n\nCommand: `{cmd}`\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')][:90]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            # L2 traffic in bytes and GB/s (a sector is 32 bytes): what `lts__t_bytes` would say
            if 'lts__t_sectors.sum' in hdr and 'gpu__time_duration.sum' in hdr:
                try:
                    sec = float(r[hdr.index('lts__t_sectors.sum')].replace(',', ''))
                    ti = hdr.index('gpu__time_duration.sum')
                    t = float(r[ti].replace(',', '')) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[ti]]
                    f.write(f"| L2 bytes (lts__t_sectors x 32) | {sec * 32 / 1e6:.1f} | MB |\n")
                    f.write(f"| L2 throughput | {sec * 32 / t / 1e9:.0f} | GB/s |\n")
                except (ValueError, KeyError):
                    pass
            f.write("\n")


NAMES = [("mlp_tc128_kernel<3>", "mlp128_fwd"), ("mlp_tc128_kernel<2>", "mlp128_bwd"), ("peer_reduce_adam", "exchange"),
         ("sample_rays", "sample_rays"), ("compact_kernel", "compact"), ("hash_fwd_kernel", "hash_fwd"),
         ("hash_bwd_kernel", "hash_bwd"), ("mlp_tc_kernel<3>", "mlp_fwd"), ("mlp_tc_kernel<0>", "mlp_fwd_plain"), ("mlp_tc_kernel<2>", "mlp_bwd"),
         ("mlp_tc_kernel<1>", "mlp_bwd_frozen"), ("composite_fwd_kernel", "composite_fwd"),
         ("composite_bwd_kernel", "composite_bwd"), ("adam_kernel<1>", "adam_table")]


def traffic(src, dst, title, cmd):
    """per-kernel DRAM bytes per launch (first captured launch of each kernel) as JSON, read by bench.py"""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        for pat, key in NAMES:
            if pat in r[ki]:      # the LAST captured launch of each kernel (warm caches, trained-on table)
                tot = 0.0
                for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                    i = hdr.index(m)
                    tot += float(r[i].replace(',', '')) * scale[units[i]]
                i = hdr.index('gpu__time_duration.sum')
                t = float(r[i].replace(',', '')) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[i]]
                res[key] = {"dram_bytes_per_launch": int(tot), "gpu_time_ms": t}
    with open(dst, 'w') as f:
        json.dump({"source": title, "command": cmd, "kernels": res}, f, indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:6])
