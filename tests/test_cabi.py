"""The C-ABI driven from plain C (tests/cabi/test_group.c): a device group of two engines, the wall-clock budget of
tnb_run_timed, the exchange step tnb_group_get_best -- no Python on the path.  On a CPU box the program links the
kernel-logic emulation build (tests/emu); with -m gpu it links the CUDA library and uses two GPUs when present."""
import os
import subprocess

import pytest

from helpers import ROOT

SRC = os.path.join(ROOT, 'tests', 'cabi', 'test_group.c')


def _build_and_run(libdir, libname, tmp_path, args=()):
    exe = str(tmp_path / 'test_group')
    subprocess.check_call(['gcc', '-O1', '-Wall', '-I', os.path.join(ROOT, 'include'), SRC, '-o', exe, '-L', libdir,
                           '-l' + libname, '-Wl,-rpath,' + libdir, '-lm'])
    out = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith('ok:'), out.stdout
    return out.stdout


def test_group_api_from_c_on_the_emulation_build(tmp_path):
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'tests', 'emu')], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    _build_and_run(os.path.join(ROOT, 'tests', 'emu'), 'tnb_emu', tmp_path)


@pytest.mark.gpu
def test_group_api_from_c_on_the_gpu(tmp_path):
    import torch
    devs = (0, 1) if torch.cuda.device_count() >= 2 else (0, 0)
    print(_build_and_run(os.path.join(ROOT, 'tnco_b200'), 'tnco_b200', tmp_path, devs))
