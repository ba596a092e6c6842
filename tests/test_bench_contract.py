"""CPU checks of bench.py's contract pieces that do not need a GPU: the per-kernel work model behind `roofline.achieved`,
the committed ncu traffic table, the peaks fallback, and the JSON line of the reference (CPU) arm on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def test_kernel_work_model_matches_design_table():
    c = dict(bench.CFG2)
    T, B = 26_400, 1024
    U = T * c['embed_dim'] * 4
    w = bench.kernel_work('sasrec_fwd_fused', T, B, c)
    assert w == dict(bound='hbm', work=(2 + 10 * c['layer_num']) * U, unit='GB/s')
    assert bench.kernel_work('attn_bwd', T, B, c)['work'] == 7 * U
    assert bench.kernel_work('adam_table', T, B, c)['work'] == 24 * c['num_items'] * c['embed_dim']
    g = bench.kernel_work('wgrad_tc', T, B, c)
    D, F = c['embed_dim'], c['hidden_size']
    assert g['bound'] == 'tensor' and g['work'] == 2 * T * (3 * D * D + D * D + 2 * D * F)
    assert bench.kernel_work('no_such_kernel', T, B, c) is None


def test_ncu_traffic_table_is_committed_and_names_profiled_kernels():
    t = bench.ncu_traffic()
    assert {'attn_bwd_tc', 'sasrec_fwd_fused', 'sasrec_bwd_ffn_fused', 'adam_table'} <= set(t)
    assert all(isinstance(v, (int, float)) and v > 0 for v in t.values())
    src = json.load(open(os.path.join(REPO, 'profiles', 'r2_ncu_traffic.json')))['source']
    assert 'ncu --set full' in src


def test_peaks_have_hbm_and_tensor_denominators():
    p = bench.peaks()
    assert p['hbm'] > 1000 and p['tensor'] > 100 and p['source'] in ('measured', 'fallback')


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` = the oracle port on the host cores; one JSON line with the contract keys."""
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=580, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == bench.UNIT and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': bench.UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_cfg1_shapes_and_identical_workload_strings():
    """`--config cfg1`: amazon-toys sizes with the shipped data's sequence-length profile (mean 2.25); both arms print the same
    `config.workload` string (the driver compares them)."""
    import torch
    from dr4sr_b200.data.synthetic import synthetic_batch
    c = bench.CFG1
    assert (c['num_items'], c['embed_dim'], c['batch_per_gpu'], c['max_seq_len']) == (11_925, 64, 256, 50)
    b = synthetic_batch(4096, 50, c['num_items'], seed=1, with_neg=False, mean_len=c['mean_len'])
    sl = b['seqlen'].double()
    assert 2.1 < float(sl.mean()) < 2.4 and int(sl.min()) >= 1 and int(sl.max()) <= 50
    live = torch.arange(50).view(1, -1) < b['seqlen'].view(-1, 1)
    assert bool((b['in_item_id'][live] > 0).all()) and bool((b['in_item_id'][~live] == 0).all())
    assert bench.workload_label(c) == bench.workload_label(dict(c)) and 'configs[0]' in bench.workload_label(c)
    assert 'configs[1]' in bench.workload_label(bench.CFG2)
