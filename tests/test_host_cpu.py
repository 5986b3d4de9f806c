"""CPU-side checks: the C-ABI library loads and exports what include/case_b200.h declares, the host
logic (sharding, gather, module face, synthetic data) behaves, and the product refuses to run
without a GPU instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from case_rg_b200 import _lib, synthetic as syn
from case_rg_b200.distributed import shard_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'case_b200.h')).read()
    declared = set(re.findall(r'^(?:int|size_t|const char\*)\s+((?:case|gttp)_\w+)\s*\(', hdr, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.case_abi_version() == 2


def test_struct_layouts_match(lib):
    for i, st in enumerate(_lib._STRUCTS):
        assert lib.case_struct_size(i) == ctypes.sizeof(st), st.__name__


def test_bad_arguments_fail_without_touching_the_gpu(lib):
    # argument validation happens before any launch, so this is safe on a CPU-only box
    rc = lib.case_topk_rows(None, 8, 1, 8, 1, None, None, None)
    assert rc == 100001 and b'case_topk_rows' in lib.case_last_error()
    a = _lib.RowLinArgs()
    assert lib.case_row_linear(ctypes.byref(a), None) == 100001


def test_sass_is_sm100a_only():
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_sass_carries_the_blackwell_instructions_the_design_claims():
    """SASS of the shipped library: tcgen05 MMAs (UTCHMMA) with TMEM loads (LDTM), TMA tensor loads (UTMALDG) of the
    producers' GEMM operands, bulk copies (UBLKCP) of the pre-packed tiles, cluster DSMEM stores of the layer kernels -
    and no Hopper wgmma."""
    out = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'HMMA.16816'):
        assert mnemonic in out, mnemonic
    assert 'HGMMA' not in out and 'WGMMA' not in out.upper().replace('UTCHMMA', '')


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from case_rg_b200.generations import FastCaSE
    sd = syn.make_case_decoder_state(1, 300, 256)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        FastCaSE(sd)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, 'case_rg_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_module_face_state_dict_and_signature():
    import inspect
    from case_rg_b200.decoder import FastCaSEDecoder
    sd = syn.make_case_decoder_state(3, 400, 256)
    m = FastCaSEDecoder(2, 4, 8, 400, 256)
    assert set(m.state_dict().keys()) == set(sd.keys()) and len(sd) == 163
    m.load_state_dict(sd, strict=True)
    params = list(inspect.signature(m.forward).parameters)
    assert params == ['encode_memories', 'BOS', 'UNK', 'source_map', 'groundtruth_index',
                      'additional_decoder_feature', 'encode_weights', 'encode_masks', 'init_decoder_state',
                      'max_target_length']
    with pytest.raises(ValueError):
        FastCaSEDecoder(2, 6, 8, 400, 256)


@pytest.mark.skipif(not os.path.isdir('/root/reference/CaSE'), reason='reference tree only exists in the build box')
def test_signature_matches_reference_module():
    import inspect
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden as mg
    from case_rg_b200.decoder import FastCaSEDecoder
    ref = mg.ref_case.CaSETransformerSeqDecoder(2, 4, 8, 300, 256)
    fast = FastCaSEDecoder(2, 4, 8, 300, 256)
    assert list(inspect.signature(ref.forward).parameters) == list(inspect.signature(fast.forward).parameters)
    assert inspect.signature(ref.forward).parameters['max_target_length'].default is None
    fast.load_state_dict(ref.state_dict(), strict=True)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in fast.state_dict().items()}


@pytest.mark.skipif(not os.path.isdir('/root/reference/Masque'), reason='reference tree only exists in the build box')
def test_masque_face_matches_reference_module():
    """FastMasqueDecoder keeps MasqueTransformerSeqDecoder's state_dict keys / shapes and forward signature
    (Masque/Model.py:13-47); install_fast_decoder picks it for a Masque-shaped model."""
    import inspect
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden  # noqa: F401  (import shim for the reference)
    import Masque.Model as ref_masque
    from case_rg_b200.decoder import FastMasqueDecoder, masque_to_case_state
    ref = ref_masque.MasqueTransformerSeqDecoder(2, 4, 8, 300, 256)
    fast = FastMasqueDecoder(2, 4, 8, 300, 256)
    assert list(inspect.signature(ref.forward).parameters) == list(inspect.signature(fast.forward).parameters)
    fast.load_state_dict(ref.state_dict(), strict=True)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in fast.state_dict().items()}
    assert set(masque_to_case_state(ref.state_dict())) == set(syn.make_case_decoder_state(3, 300, 256))
    assert set(syn.make_masque_decoder_state(3, 300, 256)) == set(ref.state_dict())


def test_shard_indices_match_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler

    class D:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n
    for n in (1, 7, 8, 260, 1008):
        for world in (1, 2, 4, 8):
            for rank in range(world):
                want = list(DistributedSampler(D(n), num_replicas=world, rank=rank, shuffle=False))
                assert shard_indices(n, rank, world) == want


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from case_rg_b200.distributed import gather_answers, shard_indices
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    n_total, T = 11, 5
    mine = shard_indices(n_total, rank, world)
    ids = torch.tensor(mine)
    toks = torch.stack([torch.arange(T) + 10 * i for i in mine])[:, :4]      # ragged: one column short
    all_ids, all_toks = gather_answers(ids, toks, T, n_total)
    q.put((rank, all_ids.tolist(), all_toks.tolist()))
    dist.destroy_process_group()


def test_gather_answers_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    for rank, ids, toks in res:
        assert ids == list(range(11))
        for i, row in enumerate(toks):
            assert row == [10 * i, 10 * i + 1, 10 * i + 2, 10 * i + 3, 0]


def test_synthetic_inputs_follow_dataset_layout():
    inp = syn.make_case_inputs(5, 4, 12, 3, 16, 1000, 256)
    assert inp.source_map.shape == (4, 12 + 3 * 16)
    assert torch.equal(inp.source_map[:, :12], inp.query.view(4, -1))
    assert torch.equal(inp.source_map[:, 12:], inp.passage.view(4, -1))
    assert (inp.query[:, 0, 0] == syn.CLS).all() and (inp.passage[:, :, 0] == syn.CLS).all()
    assert torch.allclose(inp.prior_p.view(4, -1).sum(1), torch.ones(4), atol=1e-5)
    assert float((inp.mem_p * (~inp.passage.ne(0)).unsqueeze(-1)).abs().max()) == 0.0
    a, b = syn.make_case_inputs(5, 2, 12, 3, 16, 1000, 256), syn.make_case_inputs(5, 2, 12, 3, 16, 1000, 256)
    assert torch.equal(a.mem_q, b.mem_q) and torch.equal(a.passage, b.passage)


def test_torch_custom_op_layer_registers_and_has_no_cpu_kernels():
    """TORCH_LIBRARY(case_b200, ...): the op library loads next to the C ABI, every op named in csrc/torch_ops.cpp is
    registered with the dispatcher, and the compute ops have CUDA kernels only - a CPU tensor is refused by the
    dispatcher (no compute runs here: there is no GPU on this box)."""
    ops = _lib.load_torch_ops()
    for name in ('topk_rows', 'copy_scatter_', 'softmax_mix', 'vocab_gemm', 'cross_attn_part', 'additive_attn_gate',
                 'decode_step', 'gttp_step'):
        assert hasattr(ops, name), name
    with pytest.raises(NotImplementedError):
        ops.topk_rows(torch.zeros(2, 8), 8, 1)
    with pytest.raises(NotImplementedError):
        ops.softmax_mix(torch.zeros(2, 8), torch.zeros(2, 4), 8, False)
    with pytest.raises(RuntimeError):          # the step ops check their argument blob before anything is launched
        ops.decode_step(torch.zeros(16, dtype=torch.uint8), 0)
    schema = torch.ops.case_b200.copy_scatter_.default._schema
    assert schema.arguments[0].alias_info is not None and schema.arguments[0].alias_info.is_write


@pytest.mark.timeout(300)
def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the arm the driver runs next to the B200 arm) on a tiny budget: exactly one JSON line on
    stdout carrying the contract's keys, produced by the unmodified reference when its sources are at hand (else by the
    oracle port), without touching a GPU."""
    import json
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--config', 'c1', '--steps', '1',
                          '--warmup', '0', '--cpu-budget', '6'], capture_output=True, text=True, env=env, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype',
              'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'answer_tokens_per_s' and d['unit'] == 'tokens/s'
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 0
