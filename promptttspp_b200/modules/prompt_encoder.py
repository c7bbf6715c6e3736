"""Prompt encoder: external BERT sentence embedding + the adaptor MLP weights.

Reference: promptttspp/modules/prompt_encoder.py:22-56.  ``BertWrapper`` runs BERT
natively (modules/bert.py) under HF's parameter names so the
checkpoint keys load; any callable ``(prompts, device) -> [B, in_channels]`` can
take its place (``bert=...``) -- benchmarks use a fixed-embedding provider, as
BASELINE.json's configs prescribe.
The adaptor MLP + L2-normalise + style MDN run in csrc/acoustic.cu.
"""
from typing import List

import torch
from torch import nn


class BertWrapper(nn.Module):
    """CLS embedding of BERT (prompt_encoder.py:22-38).  `model` is the native encoder (modules/bert.py) under HF's
    state_dict names; the tokenizer is HF's (host string work) when its vocabulary is available.  `forward` accepts the
    reference's list of strings, or already tokenised `(input_ids, attention_mask)` tensors."""

    def __init__(self, class_name="bert-base-uncased", **bert_config):
        super().__init__()
        from .bert import NativeBert

        self.class_name = class_name
        self.model = NativeBert(**bert_config)
        self.tokenizer = None

    def _tokenizer(self):
        if self.tokenizer is None:
            from transformers import BertTokenizer

            self.tokenizer = BertTokenizer.from_pretrained(self.class_name)  # needs the vocabulary file on disk
        return self.tokenizer

    @torch.no_grad()
    def forward(self, prompts, device) -> torch.Tensor:
        if isinstance(prompts, (tuple, list)) and len(prompts) == 2 and isinstance(prompts[0], torch.Tensor):
            ids, mask = prompts
            return self.model(ids.to(device), mask.to(device))[:, 0, :]
        # strings: the CLS vector depends only on the string (padding is masked), so serving the same style prompts
        # again costs a dictionary lookup; the cache dies with the packed weights it was computed from
        prompts = list(prompts)
        sig = self.model.weights_signature()
        if getattr(self, "_cache_sig", None) != sig or len(self._cache) > 4096:
            self._cache, self._cache_sig = {}, sig
        todo = [p for p in dict.fromkeys(prompts) if p not in self._cache]
        if todo:
            enc = self._tokenizer()(todo, padding=True, return_tensors="pt")
            cls = self.model(enc["input_ids"].to(device), enc["attention_mask"].to(device))[:, 0, :]
            for p, v in zip(todo, cls):
                self._cache[p] = v
        return torch.stack([self._cache[p].to(device) for p in prompts])


class FixedPromptEmbedding(nn.Module):
    """Parameter-free stand-in for BERT: returns rows of a fixed [N, D] table.

    ``prompts`` may be a list of strings (hashed onto rows) or a tensor of CLS
    embeddings that is passed through unchanged.
    """

    def __init__(self, table: torch.Tensor):
        super().__init__()
        self.table = table

    def forward(self, prompts, device):
        if isinstance(prompts, torch.Tensor):
            return prompts.to(device=device, dtype=torch.float32)
        rows = [i % self.table.shape[0] for i in range(len(prompts))]
        return self.table[rows].to(device)


class PromptEncoder(nn.Module):
    def __init__(self, model_name, in_channels, mid_channels, out_channels, bert=None):
        super().__init__()
        self.bert = BertWrapper(model_name) if bert is None else bert
        self.adaptor = nn.Sequential(
            nn.Linear(in_channels, mid_channels), nn.ReLU(inplace=True),
            nn.Linear(mid_channels, mid_channels), nn.ReLU(inplace=True),
            nn.Linear(mid_channels, out_channels),
        )

    def sentence_embedding(self, prompts, device) -> torch.Tensor:
        """[B, in_channels] sentence embeddings of `prompts`: a list of strings (BERT), a pair of token tensors, or an
        already computed float tensor [B, in_channels] (e.g. cached CLS vectors of recurring prompts), passed through."""
        if isinstance(prompts, torch.Tensor) and prompts.is_floating_point():
            return prompts.to(device=device, dtype=torch.float32)
        if isinstance(prompts, str):
            prompts = [prompts]
        return self.bert(prompts, device)
