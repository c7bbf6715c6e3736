"""Prompt encoder: external BERT sentence embedding + the adaptor MLP weights.

Reference: promptttspp/modules/prompt_encoder.py:22-56.  BERT-base itself is a
third-party model (HF transformers) and is SURVEY.md row f2 ("next"): this
class keeps the reference's ``bert`` sub-module so its checkpoint keys load,
and accepts any callable ``(prompts, device) -> [B, in_channels]`` in its
place (``bert=...``) -- benchmarks and tests use a fixed-embedding provider.
The adaptor MLP + L2-normalise + style MDN run in csrc/acoustic.cu.
"""
from typing import List

import torch
from torch import nn


class BertWrapper(nn.Module):
    """CLS embedding of a HF BertModel (prompt_encoder.py:22-38)."""

    def __init__(self, class_name="bert-base-uncased"):
        super().__init__()
        from transformers import BertModel, BertTokenizer

        self.model = BertModel.from_pretrained(class_name)
        self.tokenizer = BertTokenizer.from_pretrained(class_name)

    @torch.no_grad()
    def forward(self, prompts: List[str], device) -> torch.Tensor:
        inputs = self.tokenizer(prompts, padding=True, return_tensors="pt").to(device)
        return self.model(**inputs).last_hidden_state[:, 0, :]


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
        if isinstance(prompts, str):
            prompts = [prompts]
        return self.bert(prompts, device)
