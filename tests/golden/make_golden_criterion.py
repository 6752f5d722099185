"""Generate tests/golden/criterion_*.npz by running the UNMODIFIED reference criterion
(/root/reference/DASpeech/criterions/nat_dag_loss.py: NATDAGLoss.forward, its glat_function closure and
_compute_dag_loss) on CPU with the criterion's own --torch-dag-* switches.

The criterion imports fairseq (absent offline: omegaconf / hydra missing), so the four names it needs
(`fairseq.metrics`, `fairseq.utils`, `FairseqCriterion`, `register_criterion`) are provided by a throw-away stub in
sys.modules, the plugin package is assembled by file path, and the model / task are fakes that hold fixed logits and
transitions (the criterion only drives the operators; the network is irrelevant here).  Run in the build container only:

    python tests/golden/make_golden_criterion.py

tests/test_criterion_mirror.py then checks daspeech_b200.criterions (torch flavour on CPU, fused flavour on the GPU)
against these vectors.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (input generator only)

REFPKG = "/root/reference/DASpeech"
PAD = 1


def load_reference_criterion():
    fs = types.ModuleType("fairseq")
    fs.metrics = types.ModuleType("fairseq.metrics")
    fs.utils = types.ModuleType("fairseq.utils")
    fs.utils.item = lambda t: t.item() if hasattr(t, "item") else t
    fs.utils.log_softmax = lambda x, dim, onnx_trace=False: torch.log_softmax(x, dim=dim, dtype=torch.float32)
    crit = types.ModuleType("fairseq.criterions")

    class FairseqCriterion(torch.nn.Module):
        def __init__(self, task):
            super().__init__()
            self.task = task

    crit.FairseqCriterion = FairseqCriterion
    crit.register_criterion = lambda name: (lambda cls: cls)
    fs.criterions = crit
    sys.modules.update({"fairseq": fs, "fairseq.metrics": fs.metrics, "fairseq.utils": fs.utils, "fairseq.criterions": crit})

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    pkg("DASpeech", REFPKG)
    ops = load("DASpeech.custom_ops.dag_loss", os.path.join(REFPKG, "custom_ops", "dag_loss.py"))
    cops = pkg("DASpeech.custom_ops", os.path.join(REFPKG, "custom_ops"))
    for n in ("dag_loss", "dag_loss_with_alpha_beta", "dag_best_alignment", "dag_logsoftmax_gather_inplace", "torch_dag_loss",
              "torch_dag_best_alignment", "torch_dag_logsoftmax_gather_inplace", "logsumexp_keepdim"):
        setattr(cops, n, getattr(ops, n))
    pkg("DASpeech.criterions", os.path.join(REFPKG, "criterions"))
    load("DASpeech.criterions.utilities", os.path.join(REFPKG, "criterions", "utilities.py"))
    return load("DASpeech.criterions.nat_dag_loss", os.path.join(REFPKG, "criterions", "nat_dag_loss.py"))


class FakeModel(torch.nn.Module):
    """Holds the decoder outputs; forward() follows S2TConformerDAGModel.forward (models/s2t_conformer_dag.py:236-266)
    with the network replaced by the stored tensors."""

    def __init__(self, logits, links, prev_tokens):
        super().__init__()
        self.logits = torch.nn.Parameter(logits)
        self.links_p = torch.nn.Parameter(links)
        self.prev_tokens = prev_tokens
        self.pad = PAD
        self.args = types.SimpleNamespace(max_transition_length=99999)
        self.seen = {}

    def initialize_output_tokens_by_tokens(self, src_tokens, src_lengths):
        return self.prev_tokens

    def restore_valid_links(self, links):
        bsz, prelen, translen = links.shape
        idx = torch.arange(prelen).unsqueeze(1) + torch.arange(translen).unsqueeze(0) + 1
        idx = idx.masked_fill(idx >= prelen, prelen)
        res = torch.full((bsz, prelen, prelen + 1), float("-inf"), dtype=links.dtype)
        res.scatter_(2, idx.unsqueeze(0).expand(bsz, -1, -1), links)
        return res[:, :, :prelen]

    def forward(self, src_tokens, src_lengths, prev_output_tokens, tgt_tokens, glat=None, glat_function=None):
        glat_info = None
        if glat and tgt_tokens is not None:
            with torch.no_grad():
                prev_output_tokens, tgt_tokens, glat_info = glat_function(self, self.logits.detach().clone(), tgt_tokens,
                                                                          prev_output_tokens, glat, links=self.links_p.detach())
        ret = {"word_ins": {"out": self.logits * 1, "tgt": tgt_tokens, "mask": tgt_tokens.ne(self.pad), "nll_loss": True},
               "links": self.links_p * 1}
        if glat_info is not None:
            ret.update(glat_info)
        self.seen = ret
        return ret


def make_case(mod, name, B, L, M, V, T, seed, glat_p, glance_strategy):
    rng = np.random.default_rng(seed)
    _, links, olen, tlen = oracle.make_lattice(B, L, M, T, seed=seed, ragged=True)
    logits = (rng.standard_normal((B, L, V)) * 2).astype(np.float32)
    tgt = rng.integers(4, V, size=(B, M)).astype(np.int64)
    for b in range(B):
        tgt[b, tlen[b]:] = PAD
    prev = np.full((B, L), 3, dtype=np.int64)      # <unk> everywhere, pad beyond the graph (s2t_conformer_dag.py:267-283)
    prev[:, 0] = 0
    for b in range(B):
        prev[b, olen[b] - 1] = 2
        prev[b, olen[b]:] = PAD
    cfg = types.SimpleNamespace(label_smoothing=0, glance_strategy=glance_strategy, glat_p=str(glat_p), no_force_emit=False,
                                torch_dag_logsoftmax_gather=True, torch_dag_best_alignment=True, torch_dag_loss=True)
    task = types.SimpleNamespace(tgt_dict=types.SimpleNamespace(pad=lambda: PAD))
    crit = mod.NATDAGLoss(cfg, task)
    model = FakeModel(torch.tensor(logits), torch.tensor(links), torch.tensor(prev))
    sample = {"net_input": {"src_tokens": torch.zeros(B, 4), "src_lengths": torch.full((B,), 4)}, "target": torch.tensor(tgt)}
    torch.manual_seed(seed)
    loss, sample_size, log = crit(model, sample)
    loss.backward()
    out = {"logits": logits, "links": links, "olen": olen, "tlen": tlen, "tgt": tgt, "prev": prev, "glat_p": np.float64(glat_p),
           "seed": np.int64(seed), "loss": loss.detach().numpy(), "grad_logits": model.logits.grad.numpy(),
           "grad_links": model.links_p.grad.numpy(), "ntokens": np.int64(int(log["ntokens"])),
           "nvalidtokens": np.int64(int(log["nvalidtokens"])), "invalid_nsentences": np.int64(int(log["invalid_nsentences"]))}
    if glat_p > 0:
        s = model.seen
        out.update({"matchmask": s["matchmask"].numpy(), "keep_word_mask": s["keep_word_mask"].numpy(),
                    "glat_prev_output_tokens": s["glat_prev_output_tokens"].numpy(), "glat_accu": np.float64(float(s["glat_accu"])),
                    "glat_keep": np.float64(float(s["glat_keep"]))})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", float(loss), {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    mod = load_reference_criterion()
    make_case(mod, "criterion_plain", 3, 40, 12, 64, 39, seed=5, glat_p=0.0, glance_strategy=None)
    make_case(mod, "criterion_glat", 3, 40, 12, 64, 39, seed=6, glat_p=0.5, glance_strategy=None)
    make_case(mod, "criterion_glat_number_random", 2, 48, 10, 32, 16, seed=7, glat_p=0.5, glance_strategy="number-random")
