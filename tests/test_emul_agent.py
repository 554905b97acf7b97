"""Kernel logic of csrc/agent_act.cuh checked on the CPU through the host-emulation build
(tests/emul): reference goldens + eager oracle.  GPU twin: tests/test_gpu_agent.py."""
import pytest

from tests import agent_checks as AC
from tests.helpers import emul_lib


@pytest.mark.parametrize("name", ["c1", "small", "c3", "c4"])
def test_select_actions_vs_reference(name):
    AC.check_mac_against_golden(name, "cpu", emul_lib())


@pytest.mark.parametrize("name", ["c1", "small", "c3", "c4"])
def test_q_params_hidden_vs_reference(name):
    AC.check_agent_outputs_against_golden(name, "cpu", emul_lib())


def test_unroll_vs_oracle_small():
    AC.check_unroll_against_oracle("cpu", emul_lib(), O=39, A=7, H=64, AH=128, Nn=3, B=15, T=3)


def test_unroll_vs_oracle_64_row_tiles():
    # the 64-row CTA tile (RT=4) the large batches use, ragged: 2 tiles, the second partly empty
    AC.check_unroll_against_oracle("cpu", emul_lib(), O=24, A=5, H=128, AH=64, Nn=2, B=35, T=2, tile_rows=64)


@pytest.mark.parametrize("tile", [8, 16])
def test_unroll_vs_oracle_small_tiles(tile):
    AC.check_unroll_against_oracle("cpu", emul_lib(), O=24, A=5, H=64, AH=64, Nn=2, B=13, T=2, tile_rows=tile)


def test_device_rng_selection():
    AC.check_device_rng_selection("cpu", emul_lib())


@pytest.mark.parametrize("H,AH,A,O,B,T", [(256, 128, 5, 24, 20, 4), (64, 64, 7, 39, 15, 3), (128, 128, 33, 176, 12, 2)])
def test_gemm_unroll_vs_oracle(H, AH, A, O, B, T):
    """macjd_agent_unroll (the time-unrolled pass as batched layers; here on the FP32 GEMM of the host build): the
    path widths like rnn_hidden_dim = 256 take on the GPU (path 0 without the CTA-pair kernel)."""
    AC.check_unroll_against_oracle("cpu", emul_lib(), O=O, A=A, H=H, AH=AH, Nn=2, B=B, T=T, path=0)


@pytest.mark.parametrize("M,T,h0", [(1, 3, False), (11, 7, True), (8, 2, False), (299, 2, True)])
def test_recurrence_rows_kernel(M, T, h0):
    """csrc/gru_rec_rows.cuh: the recurrence of a learner-sized unroll (rows split over CTAs, weight_hh in shared memory);
    ragged last CTA, zero and given initial state, 2 and 4 rows per CTA."""
    AC.check_recurrence_rows_against_float64("cpu", emul_lib(), M=M, T=T, with_initial_state=h0)
