"""The bench contract that can be checked without a GPU: the reference arm (`bench.py --impl reference`) times the
CPU oracle port and prints ONE JSON line with the keys the driver reads; the thread-count selection honours a cgroup
CPU quota."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(built):
    env = dict(os.environ, OMP_NUM_THREADS='1')          # as exported by torchrun: the arm must not be pinned to it
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '3',
                          '--warmup', '3'], capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'agent-steps/s' and d['higher_is_better'] is True
    assert d['steps'] == 3 and d['warmup'] == 3 and d['n_gpus'] == 1 and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['value'] == d['cpu_baseline']['value'] == d['e2e']['value']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_non_zero_ranks_of_the_reference_arm_do_nothing(built):
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_cpu_quota_parser(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    real_open = open

    def fake(content):
        def _open(path, *a, **k):
            if path == '/sys/fs/cgroup/cpu.max':
                f = tmp_path / 'cpu.max'
                f.write_text(content)
                return real_open(f, *a, **k)
            return real_open(path, *a, **k)
        return _open

    monkeypatch.setattr('builtins.open', fake('1600000 100000\n'))
    assert bench._cpu_quota() == 16
    monkeypatch.setattr('builtins.open', fake('max 100000\n'))
    assert bench._cpu_quota() is None
    monkeypatch.setattr('builtins.open', fake('50000 100000\n'))
    assert bench._cpu_quota() == 1
    assert bench.alg_bytes(512) == 2144 and bench.alg_bytes(180) == 816
