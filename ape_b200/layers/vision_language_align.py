"""VisionLanguageAlign: the query x text-embedding open-vocabulary classifier.
Mirror of ape/layers/vision_language_align.py:8-52 (same constructor, same parameter names
`dot_product_projection_text`, `log_scale`, `bias_lang`, `bias0`)."""
import math

import torch
import torch.nn.functional as F
from torch import nn


class VisionLanguageAlign(nn.Module):
    def __init__(self, embed_dim, embed_dim_language, prior_prob=0.01, log_scale=0.0, clamp_dot_product=True):
        super().__init__()
        bias_value = -math.log((1 - prior_prob) / prior_prob)
        self.dot_product_projection_image = nn.Identity()
        self.dot_product_projection_text = nn.Linear(embed_dim_language, embed_dim, bias=True)
        self.log_scale = nn.Parameter(torch.Tensor([log_scale]))
        self.bias_lang = nn.Parameter(torch.zeros(embed_dim_language))
        self.bias0 = nn.Parameter(torch.Tensor([bias_value]))
        self.clamp_dot_product = clamp_dot_product

    def project_text(self, embedding, dtype):
        """Text-side half (vision_language_align.py:32-41); depends only on the vocabulary, so the
        engine computes it once per vocabulary instead of once per decoder level and image."""
        embedding = F.normalize(embedding.to(dtype), p=2, dim=-1)
        tokens = self.dot_product_projection_text(embedding / 2.0)
        bias = torch.matmul(embedding, self.bias_lang) + self.bias0
        return tokens, bias

    def forward(self, x, embedding):
        tokens, bias = self.project_text(embedding, x.dtype)
        logit = torch.matmul(x, tokens.transpose(-1, -2)) / self.log_scale.exp() + bias.unsqueeze(1)
        if self.clamp_dot_product:
            logit = torch.clamp(logit, max=50000)
            logit = torch.clamp(logit, min=-50000)
        return logit
