"""GPU parity tests of beamforming through the C-ABI against the CPU oracle."""
import numpy as np
import pytest

from oracle import cpu, numpy_ref
from seismic_bpmf_b200.beampower import beampower as bp
from tests import _synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("K,S,C,P,N,sparse,coherent", [
    (40, 5, 2, 2, 3000, True, False), (300, 8, 2, 2, 21000, False, True), (9, 3, 3, 1, 5000, False, False),
    (1000, 10, 2, 2, 9000, True, True), (1, 2, 1, 2, 700, False, False)])
@pytest.mark.parametrize("oob", ["strict", "flexible"])
def test_max_and_full_vs_oracle(K, S, C, P, N, sparse, coherent, oob):
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=K + N, sparse_sources=sparse, coherent=coherent)
    mb, am = bp.beamform(f, tau, wp, ws, device="gpu", out_of_bounds=oob, reduce="max")
    mb_w, am_w = cpu.beamform(f, tau, wp, ws, out_of_bounds=oob)
    assert mb.shape == (N,) and am.shape == (N,) and mb.dtype == np.float32
    assert np.abs(mb - mb_w).max() < 1e-5
    full = bp.beamform(f, tau, wp, ws, device="gpu", out_of_bounds=oob, reduce="none")
    assert full.shape == (K, N)
    assert np.abs(full - cpu.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none")).max() < 1e-5
    # argmax must pick a maximiser of the kernel's own beam, lowest index on ties
    assert np.array_equal(full.max(axis=0), mb)
    assert np.array_equal(full.argmax(axis=0), am)
    # where the oracle's maximiser is separated from the runner-up, indices agree exactly
    srt = np.sort(cpu.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none"), axis=0)
    clear = (srt[-1] - srt[-2] > 1e-5) if K > 1 else np.ones(N, bool)
    assert np.array_equal(am[clear], am_w[clear])


def test_float64_brute_force():
    f, tau, wp, ws = _synth.beam_case(25, 6, 2, 2, 2500, seed=3, sparse_sources=True)
    mb, _ = bp.beamform(f, tau, wp, ws)
    mb64, _ = numpy_ref.beamform(f, tau, wp, ws)
    assert np.abs(mb - mb64).max() < TOL


def test_known_answers():
    K, S, C, P, N = 50, 6, 2, 2, 6000
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=4)
    f[:] = 0
    k0, t0 = 31, 1234
    for s in range(S):
        for p in range(P):
            f[s, p % C, t0 + tau[k0, s, p]] = 1.0
    mb, am = bp.beamform(f, tau, wp, ws)
    assert abs(mb[t0] - P * ws[k0].sum()) < 1e-6 and am[t0] == k0
    mb32, am32 = bp.beamform(f, tau.astype(np.int32), wp, ws)  # int64 vs int32 moveouts
    assert np.array_equal(mb, mb32) and np.array_equal(am, am32)
    tau2 = np.concatenate([tau, tau[k0: k0 + 1]], axis=0)  # exact tie -> lowest index
    ws2 = np.concatenate([ws, ws[k0: k0 + 1]], axis=0)
    assert bp.beamform(f, tau2, wp, ws2)[1][t0] == k0
    # all-zero features: (0.0, 0) everywhere
    mb0, am0 = bp.beamform(np.zeros_like(f), tau, wp, ws)
    assert np.all(mb0 == 0) and np.all(am0 == 0)


def test_source_shard_invariance_and_offsets():
    f, tau, wp, ws = _synth.beam_case(101, 5, 2, 2, 4000, seed=5)
    mb, am = bp.beamform(f, tau, wp, ws)
    parts = [(0, 40), (40, 41), (41, 101)]
    best = np.full(4000, -np.inf, np.float32)
    arg = np.zeros(4000, np.int32)
    for lo, hi in parts:
        m, a = bp.beamform(f, tau[lo:hi], wp, ws[lo:hi], source_offset=lo)
        upd = (m > best) | ((m == best) & (a < arg))
        best[upd], arg[upd] = m[upd], a[upd]
    assert np.array_equal(best, mb) and np.array_equal(arg, am)


def test_wide_moveout_spread_restages():
    """Random (incoherent) moveouts wider than the staged row segment exercise the restaging path."""
    f, tau, wp, ws = _synth.beam_case(64, 4, 2, 2, 30000, seed=6, tau_max=6000)
    mb, am = bp.beamform(f, tau, wp, ws)
    mb_w, am_w = cpu.beamform(f, tau, wp, ws)
    assert np.abs(mb - mb_w).max() < 1e-5


def test_prestack():
    f, tau, wp, ws = _synth.beam_case(3, 4, 3, 2, 1000, seed=7)
    wp = np.random.default_rng(0).random(wp.shape).astype(np.float32)
    U = bp.prestack_traces(f, wp)
    assert np.abs(U - np.einsum("scp,scn->spn", wp, f)).max() < 1e-5


def test_grouped_kernel_matches_legacy_kernel_bitwise(monkeypatch):
    """The warp-specialised grouped kernel and the barrier-synchronised legacy kernel add the (station, phase)
    rows of a source in the same order: identical bits, coherent (groups of 16) and random (groups of 1)."""
    for coherent, sparse in ((True, False), (False, True), (True, True)):
        f, tau, wp, ws = _synth.beam_case(333, 7, 2, 2, 6000, seed=11, sparse_sources=sparse, coherent=coherent)
        monkeypatch.delenv("BPMF_BEAM_KERNEL", raising=False)
        mb, am = bp.beamform(f, tau, wp, ws)
        full = bp.beamform(f, tau, wp, ws, reduce="none")
        monkeypatch.setenv("BPMF_BEAM_KERNEL", "legacy")
        mb_l, am_l = bp.beamform(f, tau, wp, ws)
        full_l = bp.beamform(f, tau, wp, ws, reduce="none")
        monkeypatch.delenv("BPMF_BEAM_KERNEL", raising=False)
        assert np.array_equal(mb, mb_l) and np.array_equal(am, am_l) and np.array_equal(full, full_l)


@pytest.mark.parametrize("oob", ["strict", "flexible"])
def test_negative_and_out_of_range_moveouts(oob):
    """Moveouts below zero, beyond the end of the record, and a source whose every arrival is out of range."""
    K, S, C, P, N = 70, 4, 2, 2, 5000
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=21)
    rng = np.random.default_rng(5)
    tau = tau + rng.integers(-700, 300, size=(K, 1, 1))
    tau[7] += 2 * N          # entirely beyond the record
    tau[8, 1, :] = -3 * N    # one station far before the record
    tau[40:44] += N - 600    # windows that straddle the end
    mb, am = bp.beamform(f, tau, wp, ws, out_of_bounds=oob)
    full = bp.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none")
    want = cpu.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none")
    assert np.abs(full - want).max() < 1e-5
    assert np.array_equal(full.max(axis=0), mb) and np.array_equal(full.argmax(axis=0), am)


def test_unused_station_garbage_is_ignored():
    """NaN features and absurd moveouts on stations whose source weight is zero never reach the beam."""
    K, S, C, P, N = 48, 6, 2, 2, 4000
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=31, coherent=True)
    ws[:, 2] = 0.0
    ws[::3, 4] = 0.0
    ws /= ws.sum(axis=1, keepdims=True)
    want = cpu.beamform(f, tau, wp, ws, reduce="none")
    f2, tau2 = f.copy(), tau.copy()
    f2[2] = np.nan
    tau2[:, 2, :] = 2 ** 30
    tau2[::3, 4, :] = -(2 ** 30)
    got = bp.beamform(f2, tau2, wp, ws, reduce="none")
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() < 1e-5


@pytest.mark.parametrize("K,S,C,P,N", [(37, 40, 3, 3, 2500), (21, 200, 2, 2, 1800), (5, 1, 1, 1, 900)])
def test_station_counts_beyond_one_warp_and_legacy_fallback(K, S, C, P, N):
    """More (station, phase) rows than lanes (the pack kernel and the producer loop over them), and a network
    large enough (200 stations x 2 phases) that the grouped kernel's parameter buffers and running-max arrays no
    longer fit the shared memory of one SM for reduce="max": the barrier-synchronised kernel takes over there, while
    reduce="none" still runs the grouped kernel.  Same answers either way."""
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=K + S, coherent=True, sparse_sources=(S == 40))
    mb, am = bp.beamform(f, tau, wp, ws)
    full = bp.beamform(f, tau, wp, ws, reduce="none")
    want = cpu.beamform(f, tau, wp, ws, reduce="none")
    assert np.abs(full - want).max() < 2e-5 * max(1.0, float(np.abs(want).max()))
    assert np.array_equal(full.max(axis=0), mb) and np.array_equal(full.argmax(axis=0), am)
