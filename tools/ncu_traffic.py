#!/usr/bin/env python
"""Turn `ncu --page raw --csv` exports (tools/ncu_capture.sh) into the committed evidence:
  profiles/r2_ncu_traffic.json : dram read + write bytes per launch per kernel (what bench.py reports as roofline.traffic)
  profiles/r2_ncu_summary.md   : per kernel: duration, DRAM bytes and % of peak, tensor-pipe %, issue-active %, registers, smem
usage: ncu_traffic.py out_prefix raw1.csv [raw2.csv ...]"""
import collections
import csv
import json
import re
import sys

# kernel-name fragment -> the profile name bench.py / DESIGN.md use
NAMES = [('sasrec_fwd_fused_kernel', 'sasrec_fwd_fused'), ('sasrec_bwd_ffn_fused_kernel', 'sasrec_bwd_ffn_fused'),
         ('attn_bwd_tc2_kernel', 'attn_bwd_tc'), ('wgrad_tc_kernel', 'wgrad_tc'), ('gemm_tc_kernel', 'gemm_bwd_dx'),
         ('adam_kernel', 'adam'), ('table_grad_kernel', 'table_grad_scatter'), ('score_bce_kernel', 'score_bce'),
         ('prep_scan_kernel', 'prep_batch'), ('pos_grad_kernel', 'pos_grad'), ('colsum_kernel', 'colsum'),
         ('reduce_segments_kernel', 'reduce_partials'), ('fused_tiles_kernel', 'fused_tiles'), ('weight_image_kernel', 'weight_images'),
         ('neg_sample_kernel', 'neg_sample'), ('sum_kernel', 'loss_sum'), ('scale_grads_kernel', 'scale_grads'),
         ('logits_topk_kernel', 'topk_logits_tc'), ('select_kernel', 'topk_select'), ('table_image_kernel', 'topk_table_images'),
         ('gru_fwd_tc_kernel', 'gru_recurrence_fwd_tc'), ('gru_fwd_kernel', 'gru_recurrence_fwd'), ('gru_bwd_kernel', 'gru_recurrence_bwd'),
         ('gru_order_kernel', 'gru_order'), ('filter_fwd_kernel', 'fmlp_filter_fwd'), ('filter_bwd_dx_kernel', 'fmlp_filter_bwd_dx'),
         ('filter_bwd_dh_kernel', 'fmlp_filter_bwd_dh'), ('filter_wgrad_kernel', 'fmlp_filter_wgrad'), ('filter_taps_kernel', 'fmlp_filter_taps'),
         ('embed_dense_kernel', 'fmlp_embed'), ('ln_fwd_kernel', 'ln_fwd'), ('ln_bwd_kernel', 'ln_bwd'), ('embed_fwd_kernel', 'embed_fwd')]
COLS = {'gpu__time_duration.sum': 'dur', 'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
        'sm__inst_issued.avg.pct_of_peak_sustained_active': 'issue_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_pct',
        'launch__registers_per_thread': 'regs', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'launch__shared_mem_per_block_dynamic': 'smem_dyn', 'launch__shared_mem_per_block_static': 'smem_static'}
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Kibyte': 1024.0, 'Mibyte': 1024.0 ** 2}


def profile_name(kernel):
    for frag, name in NAMES:
        if frag in kernel:
            if name == 'adam':
                return name
            return name
    return None


def load(path):
    rows = list(csv.reader(open(path, newline='')))
    hdr_i = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    out = []
    for r in rows[hdr_i + 2:]:
        if len(r) != len(hdr):
            continue
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for col, key in COLS.items():
            if col in hdr:
                j = hdr.index(col)
                try:
                    v = float(r[j].replace(',', ''))
                except ValueError:
                    continue
                d[key] = v * SCALE.get(units[j], 1.0)
        out.append(d)
    return out


def main(prefix, paths):
    launches = [d for p in paths for d in load(p)]
    agg = collections.OrderedDict()
    for d in launches:
        name = profile_name(d['kernel'])
        if name is None:
            continue
        if name == 'adam':
            name = 'adam_table' if d.get('rd', 0) > 50e6 else 'adam_small'
        if name == 'gemm_bwd_dx' and 'gemm_tc_kernel' in d['kernel']:
            name = 'gemm_tc'
        agg.setdefault(name, []).append(d)
    traffic, lines = {}, []
    lines.append('| kernel | launches captured | avg us | dram read MB | dram write MB | dram % of peak | tensor pipe % | issue active % | regs | grid x block | smem KB |')
    lines.append('|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|')
    for name, ds in sorted(agg.items(), key=lambda kv: -sum(x.get('dur', 0) for x in kv[1])):
        n = len(ds)
        avg = lambda k: sum(x.get(k, 0.0) for x in ds) / n
        traffic[name] = int(avg('rd') + avg('wr'))
        lines.append(f"| `{name}` | {n} | {avg('dur'):.1f} | {avg('rd') / 1e6:.1f} | {avg('wr') / 1e6:.1f} | {avg('dram_pct'):.1f} | {avg('tensor_pct'):.1f} | "
                     f"{avg('issue_pct'):.1f} | {int(avg('regs'))} | {int(avg('grid'))} x {int(avg('block'))} | {(avg('smem_dyn') + avg('smem_static')) / 1024:.1f} |")
    json.dump({'source': 'ncu --set full --clock-control none (tools/ncu_capture.sh) over a cfg-2 SASRec step and the eval top-k batch of this commit; '
                         'dram__bytes_read.sum + dram__bytes_write.sum per launch, cold cache (ncu flushes L2 between replays); generated by tools/ncu_traffic.py',
               'bytes_per_launch': traffic}, open(prefix + '_traffic.json', 'w'), indent=1)
    open(prefix + '_summary.md', 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
