"""Kernel LOGIC on CPU: tests/emu builds the very same kernel sources (tnb_kernels.h) with one lane per chain
and host memory as the "device", so the sweep / slicer / init logic can be checked against the reference's
golden vectors on machines without a GPU.  This is test infrastructure -- the tnco_b200 package never loads
the emulation library; the real parity tests are tests/test_gpu_parity.py (-m gpu)."""
import ctypes
import os
import subprocess

import pytest

import test_gpu_parity as G

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu_lib():
    subprocess.check_call(['make', '-C', os.path.join(HERE, 'emu')], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    from tnco_b200 import _lib
    return _lib.bind(ctypes.CDLL(os.path.join(HERE, 'emu', 'libtnb_emu.so')))


@pytest.fixture()
def emu(emu_lib, monkeypatch):
    from tnco_b200 import _lib
    monkeypatch.setattr(_lib, '_LIB', emu_lib)
    yield


@pytest.mark.parametrize('name', G.CASES)
def test_emu_mt19937_matches_reference_golden(emu, name):
    G.test_mt19937_mode_matches_reference_golden.__wrapped__(name) if hasattr(
        G.test_mt19937_mode_matches_reference_golden, '__wrapped__') else G.test_mt19937_mode_matches_reference_golden(name)


@pytest.mark.parametrize('name', ['reg64_inf', 'reg100_fw30', 'reg300_inf'])
def test_emu_replay(emu, name):
    G.test_replay_of_recorded_draw_stream_is_bit_exact(name)


@pytest.mark.parametrize('n,dim', [(8, 2), (64, 2), (64, 3), (300, 2)])
def test_emu_eval_cost(emu, n, dim):
    G.test_eval_cost_matches_oracle(n, dim)


def test_emu_philox(emu):
    G.test_philox_chains_are_valid_and_deterministic()


@pytest.mark.parametrize('n,max_width', [(64, 10), (150, 22)])
def test_emu_philox_finite(emu, n, max_width):
    G.test_philox_finite_width_chains_are_valid(n, max_width, None)


def test_emu_philox_skip_slices(emu):
    G.test_philox_skip_slices_are_never_sliced()


@pytest.mark.parametrize('dim', [3])
def test_emu_philox_uniform_dim3_finite_width(emu, dim):
    G.test_philox_uniform_dimension_other_than_two_with_max_width(dim)


@pytest.mark.parametrize('n,method', [(2, 0), (3, 1), (64, 0), (64, 1), (180, 0)])
def test_emu_generated_trees(emu, n, method):
    G.test_device_generated_trees_are_valid(n, method)


def test_emu_generated_trees_disconnected(emu):
    G.test_device_generated_trees_reject_disconnected_network()


@pytest.mark.parametrize('max_width', [None, 20])
def test_emu_split_layout(emu, max_width):
    G.test_split_layout_gives_identical_results(max_width)


@pytest.mark.parametrize('n,max_width,n_sweeps', [(64, None, 1000), (100, 14, 1000)])
def test_emu_statistics(emu, n, max_width, n_sweeps):
    import test_gpu_statistics as S
    S.test_best_cost_distribution_is_no_worse_than_the_reference(n, max_width, n_sweeps)


@pytest.mark.parametrize('n,max_width', [(40, None), (90, 18)])
def test_emu_hyper_philox(emu, n, max_width):
    G.test_philox_hyper_index_networks_are_valid(n, max_width)


@pytest.mark.parametrize('name', ['hyper64_inf', 'hyper64_fw40'])
def test_emu_replay_hyper(emu, name):
    G.test_replay_of_recorded_draw_stream_is_bit_exact(name)


@pytest.mark.parametrize('n,max_width,hyper', [(80, None, False), (80, 30, False), (60, 26, True)])
def test_emu_per_index_dims(emu, n, max_width, hyper):
    G.test_philox_per_index_dims(n, max_width, hyper)
    G.test_per_index_dims_must_be_powers_of_two()


@pytest.mark.parametrize('name', ['dims64_fw45', 'dimshyper48_fw50'])
def test_emu_replay_dims(emu, name):
    G.test_replay_of_recorded_draw_stream_is_bit_exact(name)


def test_emu_packed_trees_and_cache(emu):
    G.test_packed_tree_readback_and_cached_engine()


def test_emu_minima_exact_on_slow_ramp(emu):
    G.test_minima_stay_exact_when_chains_wander_at_small_beta()


@pytest.mark.parametrize('dim', [2, 3])
def test_emu_sparse_philox(emu, dim):
    G.test_philox_sparse_index_chains_are_valid(dim)


@pytest.mark.parametrize('name', ['sparse64_inf', 'sparse100_fw30', 'sparsehyper64_fw40'])
def test_emu_replay_sparse(emu, name):
    G.test_replay_of_recorded_draw_stream_is_bit_exact(name)


def test_emu_mode_and_resume_errors(emu):
    G.test_mode_and_resume_argument_errors()


@pytest.mark.parametrize('n,max_width,sparse', [(60, None, False), (60, 16.0, False), (60, 16.0, True)])
def test_emu_general_dims_philox(emu, n, max_width, sparse):
    G.test_philox_general_per_index_dims(n, max_width, sparse)


@pytest.mark.parametrize('name,n_sweeps', [('C1', 1500), ('C2', 500), ('C4', 200), ('reg200_fw', 300)])
def test_emu_production_kernel_replay(emu, name, n_sweeps):
    """Same kernel source, one lane per chain: the recorded decisions replayed through the oracle (see
    tests/test_gpu_philox_replay.py; the GPU run is the parity test proper)."""
    import test_gpu_philox_replay as R
    net, mw = R.NETS[name]
    R.run_and_replay(net, mw, n_sweeps, chains=2)


def test_emu_production_kernel_replay_low_beta(emu):
    import test_gpu_philox_replay as R
    R.test_production_kernel_replay_low_beta_and_every_sweep_reslice()


@pytest.mark.parametrize('case,chains', [('c2_1e4', 96), ('c4_1e4', 64)])
def test_emu_equal_sweep_distribution(emu, case, chains):
    import test_gpu_statistics as S
    S.test_equal_sweep_distribution_on_the_benchmarked_networks(case, chains)


@pytest.mark.parametrize('n,method', [(12, 0), (48, 0), (48, 1)])
def test_emu_generated_trees_hyper(emu, n, method):
    G.test_device_generated_trees_for_hyper_index_networks(n, method)
