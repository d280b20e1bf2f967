"""Hardware probe of the TS form of tcgen05.mma (A operand in tensor memory) that the round-2 tensor-core kernels
build on: TMEM layout of a TF32 A operand (row = lane, one 32-bit column per K element), tcgen05.st -> MMA ordering,
SWIZZLE_64B K-major B tiles written by ordinary stores -- and what the tensor core does with the low 13 mantissa bits
of an unrounded fp32 input (the result is recorded in gpurun_out/tcts_probe.json for DESIGN.md)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_rna(x):
    return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize('bn', [16, 32])
def test_ts_mma_exact_on_tf32_inputs_and_low_bit_behaviour(bn):
    import nlt_native as nat
    lib = nat.lib()
    dev = torch.device('cuda')
    rng = np.random.default_rng(5 + bn)
    # 1) TF32-exact inputs (small integers / 8): the product is exact in fp32 -> bit-exact check of the data path
    A = (rng.integers(-16, 17, size=(128, 16)) / 8.0).astype(np.float32)
    B = (rng.integers(-16, 17, size=(bn, 16)) / 8.0).astype(np.float32)
    out = torch.empty(128, bn, dtype=torch.float32, device=dev)
    a_d, b_d = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    nat.check(lib.nlt_debug_tcts_probe(nat.ptr(a_d), nat.ptr(b_d), bn, nat.ptr(out), nat.stream()))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), A @ B.T)
    # 2) full-mantissa fp32 A against TF32-exact B: truncation or rounding of the low 13 bits?
    A2 = rng.standard_normal((128, 16)).astype(np.float32)
    a_d = torch.from_numpy(A2).to(dev)
    nat.check(lib.nlt_debug_tcts_probe(nat.ptr(a_d), nat.ptr(b_d), bn, nat.ptr(out), nat.stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    want_trunc = _tf32_trunc(A2).astype(np.float64) @ B.T.astype(np.float64)
    want_rna = _tf32_rna(A2).astype(np.float64) @ B.T.astype(np.float64)
    want_full = A2.astype(np.float64) @ B.T.astype(np.float64)
    err = {k: float(np.abs(got - w).max()) for k, w in (('trunc', want_trunc), ('rna', want_rna), ('full', want_full))}
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'tcts_probe_bn%d.json' % bn), 'w') as f:
        json.dump(err, f)
    # whichever it is, the hardware result must be one of the two TF32 readings to fp32 accumulation accuracy
    assert min(err['trunc'], err['rna']) <= 1e-5, err
