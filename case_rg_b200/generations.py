"""Protocol face: ``greedy`` / ``beam`` with the signatures of common/Generations.py:66,112, and the
models they drive.

The reference drives an ``EncDecModel`` (GTTP/EncDecModel.py:11-42) from Python, one ``.item()``
per hypothesis per step.  Here the whole search runs on the device; the functions below only hand
the model's encoder outputs to an engine and return the same LongTensor the reference returns
(``[B, max_len]`` for greedy; for beam the best hypothesis per query without BOS, EOS kept,
zero-padded to the longest answer - ``merge1D``, common/Utils.py:366-377).

``FastCaSE`` gives the unchanged CaSE decoder the protocol the reference never implemented for it
(SURVEY.md §8c adapter); ``FastGTTP`` replaces ``GTTP.decode/generate/to_word`` + the search
(GTTP/Model.py:176-193).  Both take the *encoder-side outputs* as inputs: the encoders are outside
the hot path and stay on the reference's own code.
"""
from typing import Dict, Optional

import torch

from . import _lib as L
from .engine import CaseDecodeEngine, CaseWeights, GttpDecodeEngine, GttpWeights

PAD_WORD, BOS_WORD, UNK_WORD, EOS_WORD = '[PAD]', '[unused0]', '[UNK]', '[unused1]'


def _check_vocab(vocab2id):
    if vocab2id is None:
        return
    want = {PAD_WORD: 0, BOS_WORD: 1, EOS_WORD: 2, UNK_WORD: 100}
    for w, i in want.items():
        if vocab2id.get(w, i) != i:
            raise ValueError(f'special token {w} must have id {i} (BERT-uncased ids, common/Constants.py:1-7)')


def greedy(model, data, vocab2id=None, max_len=20, encode_outputs=None, init_decoder_states=None):
    """Generations.greedy (Generations.py:66-110) on the device."""
    _check_vocab(vocab2id)
    return model.fast_search(data, max_len, 1, L.MODE_PROTO_GREEDY, encode_outputs, init_decoder_states)


def beam(model, data, vocab2id=None, max_len=20, width=5, encode_outputs=None, init_decoder_states=None):
    """Generations.beam (Generations.py:112-190) on the device."""
    _check_vocab(vocab2id)
    return model.fast_search(data, max_len, width, L.MODE_BEAM, encode_outputs, init_decoder_states)


def beam_batches(model, batches, vocab2id=None, max_len=20, width=5):
    """``beam`` over an iterable of HOST batches, the way the predict loop feeds a model one batch after
    another (common/CumulativeTrainer.py:141-156): yields the answers of every batch, in order, as host
    LongTensors.  The host-to-device copy of batch i+1 runs on a copy stream while batch i is decoding
    (double-buffered staging, pinned host tensors copy asynchronously), so the PCIe time of the encoder
    outputs (175 MB per batch at the BASELINE shape) is hidden behind the decode."""
    _check_vocab(vocab2id)
    return model.search_batches(batches, max_len, width, L.MODE_BEAM)


def greedy_batches(model, batches, vocab2id=None, max_len=20):
    """``greedy`` over an iterable of host batches with the copy of the next batch overlapped (see beam_batches)."""
    _check_vocab(vocab2id)
    return model.search_batches(batches, max_len, 1, L.MODE_PROTO_GREEDY)


class _FastModel:
    beam_width = 1
    max_dec_len = 40

    def greedy(self, data):                       # EncDecModel.greedy / .beam (EncDecModel.py:38-42)
        return greedy(self, data, None, self.max_dec_len)

    def beam(self, data):
        return beam(self, data, None, self.max_dec_len, self.beam_width)

    def forward(self, data, method='test'):       # GTTP.forward test branch (GTTP/Model.py:204-212)
        if method != 'test':
            raise NotImplementedError('only the test-mode decode path is implemented on the device')
        return {'answer': self.greedy(data) if self.beam_width == 1 else self.beam(data)}

    __call__ = forward

    id2vocab = None

    def to_sentence(self, data, batch_indices):    # CaSE.to_sentence (CaSE/Model.py:270-271); needs self.id2vocab
        from .results import to_sentence
        if self.id2vocab is None:
            raise ValueError('set model.id2vocab before calling to_sentence')
        return to_sentence(batch_indices, self.id2vocab)

    # ---- streamed batches (beam_batches / greedy_batches)
    def search_batches(self, batches, max_len, width, mode):
        """Generator behind beam_batches / greedy_batches."""
        dev = self.weights.device
        main = torch.cuda.current_stream(dev)
        # the copy stream and the two staging sets live as long as the model: allocating 2 x 175 MB on a
        # fresh stream per call costs tens of milliseconds of cudaMalloc
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream, self._staging = torch.cuda.Stream(dev), [None, None]
        copy, staging = self._copy_stream, self._staging
        copy.wait_stream(main)              # earlier readers of the staging sets (previous call) are done
        prefilled = [None, None]            # event: the prefill that last read staging set k has finished

        def stage(k, host):
            shapes = {n: (tuple(host[n].shape), host[n].dtype) for n in self._KEYS}
            with torch.cuda.stream(copy):
                if staging[k] is None or staging[k][0] != shapes:
                    # allocated ON the copy stream: a block the allocator recycles from the main stream's pool
                    # (e.g. a prefill temporary that was freed on the host but is still being read on the
                    # device) must never be handed to a buffer the copy stream writes
                    staging[k] = (shapes, {n: torch.empty(host[n].shape, dtype=host[n].dtype, device=dev)
                                           for n in self._KEYS})
                if prefilled[k] is not None:
                    copy.wait_event(prefilled[k])
                for n in self._KEYS:
                    staging[k][1][n].copy_(host[n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return ev

        it = iter(batches)
        nxt = next(it, None)
        k = 0
        ready = stage(0, nxt) if nxt is not None else None
        pending = None                      # engine whose decode is in flight
        while nxt is not None:
            d = staging[k][1]
            # stream order: decode(i-1) -> prefill(i); the host enqueues prefill(i) while decode(i-1) runs.
            # (prefill never touches the search state the pending answers are read from; an engine that is still
            # decoding is never prefilled: engine_for returns the SAME engine for equal shapes, and its decode is
            # ahead of this prefill on the main stream)
            main.wait_event(ready)
            eng = self._prefill_staged(d, width, max_len)
            prefilled[k] = torch.cuda.Event()
            prefilled[k].record(main)
            nxt = next(it, None)
            if nxt is not None:             # the next batch starts crossing PCIe before this one decodes
                ready = stage(k ^ 1, nxt)
            if mode != L.MODE_BEAM and eng.W != 1:
                raise ValueError('greedy modes need an engine built with W == 1')
            if pending is not None:
                # answers of the previous batch: snapshot them on the device (stream-ordered after its decode), put
                # THIS batch's decode behind the snapshot right away, and only then wait for the snapshot on a side
                # stream - the device never idles while the host reads answers
                snap = (pending.state.out_tokens[:, :max_len].clone(), pending.state.best_len.clone())
                ev = torch.cuda.Event()
                ev.record(main)
                eng.launch(max_len, mode, use_graph=self.use_graph)
                yield self._read_snapshot(snap, ev, mode)
            else:
                eng.launch(max_len, mode, use_graph=self.use_graph)
            self.last_engine = pending = eng
            k ^= 1
        if pending is not None:
            yield pending._finish_tokens(max_len, mode).cpu()

    def _read_snapshot(self, snap, ev, mode):
        """Device snapshot (tokens [B, T] int32, best_len [B]) -> host int64 tokens, trimmed like _finish_tokens."""
        dev = self.weights.device
        if getattr(self, '_d2h_stream', None) is None:
            self._d2h_stream = torch.cuda.Stream(dev)
        d2h = self._d2h_stream
        d2h.wait_event(ev)
        with torch.cuda.stream(d2h):
            toks = snap[0].to('cpu', non_blocking=True)
            blen = snap[1].to('cpu', non_blocking=True)
        d2h.synchronize()
        out = toks.to(torch.int64)
        if mode == L.MODE_BEAM:             # merge1D (Utils.py:366-377): pad to the longest answer of the batch
            out = out[:, :max(int(blen.max()), 1)]
        return out


class FastCaSE(_FastModel):
    """EncDecModel-protocol driver for the CaSE decoder.

    ``data`` carries what ``ResponseGeneration.action`` hands the decoder (CaSE/Model.py:247-251):
    'mem_q' [B,1,Lq,H], 'mem_p' [B,NP,Lp,H], 'query' / 'passage' ids (masks = ids != 0),
    'prior_q', 'prior_p', 'answer_rep' [B,H], 'source_map' int64 [B,S]."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None, dtype='bf16', max_dec_len=40,
                 beam_width=1, vocab_impl: Optional[int] = None, use_graph=True, prefix='', opt: int = 0,
                 n_oov: int = 0, producers: Optional[Dict[str, torch.Tensor]] = None):
        """opt: CASE_OPT_* bits for every engine of this model (0 = the default fast path; A/B and fallback tests).
        n_oov: size of the per-query dynamic vocabulary behind the V fixed ids (``source_map`` may point into
        [V, V + n_oov); returned ids then range over the extended vocabulary)."""
        self.weights = CaseWeights(state_dict, device=device, dtype=dtype, prefix=prefix)
        self.max_dec_len, self.beam_width, self.vocab_impl, self.use_graph = max_dec_len, beam_width, vocab_impl, use_graph
        self.opt, self.n_oov = int(opt), int(n_oov)
        self._engines = {}
        # producers: state_dict of the pre-decode modules (reference key names 'query_encoder.*', 'passage_selection.*',
        # 'span_extraction.*') -> search_ids decodes from token ids, the encoder outputs never leave the device
        self.producers = None
        if producers is not None:
            from .producers import CaseProducers
            self.producers = CaseProducers(producers, device=self.weights.device)

    def search_ids(self, data, max_len=None, width=None, mode='beam'):
        """CaSE.do_test from token ids (CaSE/Model.py:313-331): ``data`` = {'query' int [B,1,Lq], 'passage' int [B,NP,Lp],
        'source_map' int [B,S]} -> {'answer': LongTensor, 'rank': [B,NP]} like the reference's forward(data, 'test')."""
        if self.producers is None:
            raise RuntimeError('FastCaSE was built without producers=...')
        max_len = self.max_dec_len if max_len is None else max_len
        width = self.beam_width if width is None else width
        p = self.producers(data['query'], data['passage'])
        d = dict(mem_q=p['mem_q'], mem_p=p['mem_p'], query=data['query'].to(self.weights.device),
                 passage=data['passage'].to(self.weights.device), prior_q=p['prior_q'], prior_p=p['prior_p'],
                 answer_rep=p['answer_rep'], source_map=data['source_map'].to(self.weights.device))
        m = {'beam': L.MODE_BEAM, 'greedy': L.MODE_PROTO_GREEDY, 'module_greedy': L.MODE_MODULE_GREEDY}[mode]
        return {'answer': self.fast_search(d, max_len, width, m), 'rank': p['rank']}

    def engine_for(self, B, W, S0, S1, T):
        key = (B, W, S0, S1, T, self.opt, self.n_oov)
        if key not in self._engines:
            if len(self._engines) >= 4:
                self._engines.clear()
            self._engines[key] = CaseDecodeEngine(self.weights, B, W, S0, S1, T, vocab_impl=self.vocab_impl,
                                                  opt=self.opt, n_oov=self.n_oov)
        return self._engines[key]

    def encode(self, data):
        return data

    def fast_search(self, data, max_len, width, mode, encode_outputs=None, init_decoder_states=None):
        d = encode_outputs if encode_outputs is not None else data
        B = d['source_map'].size(0)
        S0 = d['mem_q'].reshape(B, -1, L.H).size(1)
        S1 = d['mem_p'].reshape(B, -1, L.H).size(1)
        eng = self.engine_for(B, width, S0, S1, max_len)
        eng.prefill(d['mem_q'], d['mem_p'], d['query'].ne(0), d['passage'].ne(0), d['prior_q'], d['prior_p'],
                    d['answer_rep'], d['source_map'])
        self.last_engine = eng
        return eng.decode(max_len, mode, use_graph=self.use_graph)

    _KEYS = ('mem_q', 'mem_p', 'query', 'passage', 'prior_q', 'prior_p', 'answer_rep', 'source_map')

    def _prefill_staged(self, d, width, max_len):
        B = d['source_map'].size(0)
        S0 = d['mem_q'].reshape(B, -1, L.H).size(1)
        S1 = d['mem_p'].reshape(B, -1, L.H).size(1)
        eng = self.engine_for(B, width, S0, S1, max_len)
        eng.prefill(d['mem_q'], d['mem_p'], d['query'].ne(0), d['passage'].ne(0), d['prior_q'], d['prior_p'],
                    d['answer_rep'], d['source_map'])
        return eng

    def module_greedy(self, data, max_len):
        """The in-module loop of CaSETransformerSeqDecoder.forward (no EOS handling, Model.py:91-123)."""
        return self.fast_search(data, max_len, 1, L.MODE_MODULE_GREEDY)


class FastGTTP(_FastModel):
    """Step side of GTTP on the device.  ``data``: 'context' [B,Lc], 'background' [B,Lb],
    'background_map' int64 [B,Lb] (index form, never one-hot), and the encoder outputs
    'src_output' [B,Lc,2H], 'bg_output' [B,Lb,2H], 'init_state' [B,1,H] (GTTP/Model.py:156-174)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None, dtype='bf16', max_dec_len=40,
                 beam_width=1, vocab_impl: Optional[int] = None, use_graph=True, prefix='', opt: int = 0):
        self.weights = GttpWeights(state_dict, device=device, dtype=dtype, prefix=prefix)
        self.max_dec_len, self.beam_width, self.vocab_impl, self.use_graph = max_dec_len, beam_width, vocab_impl, use_graph
        self.opt = int(opt)
        self._engines = {}

    def engine_for(self, B, W, Lc, Lb, T):
        key = (B, W, Lc, Lb, T)
        if key not in self._engines:
            if len(self._engines) >= 4:
                self._engines.clear()
            self._engines[key] = GttpDecodeEngine(self.weights, B, W, Lc, Lb, T, vocab_impl=self.vocab_impl,
                                                  opt=self.opt)
        return self._engines[key]

    def fast_search(self, data, max_len, width, mode, encode_outputs=None, init_decoder_states=None):
        d = dict(data)
        if encode_outputs is not None:
            d.update(encode_outputs)
        if init_decoder_states is not None:
            d['init_state'] = init_decoder_states
        eng = self._prefill_staged(d, width, max_len)
        self.last_engine = eng
        return eng.decode(max_len, mode, use_graph=self.use_graph)

    _KEYS = ('context', 'background', 'background_map', 'src_output', 'bg_output', 'init_state')

    def _prefill_staged(self, d, width, max_len):
        B, Lc = d['context'].shape
        Lb = d['background'].size(1)
        eng = self.engine_for(B, width, Lc, Lb, max_len)
        eng.prefill(d['src_output'], d['bg_output'], d['context'], d['background'], d['background_map'],
                    d['init_state'])
        return eng
