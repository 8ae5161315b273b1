"""BertEncoder of MIBF-Net (mibf_net/bert.py:6-13): CLS vector of the last hidden state, on the B200 BERT engine."""
import torch

from ..bert_engine import BertEngine, qkv_groups
from ..encoder import MdhsModule, _BertFn


class BertEncoder(MdhsModule):
    def __init__(self, model_path="/data/QLI/BERT_pretain"):
        super().__init__()
        from transformers import BertModel
        self.bert = BertModel.from_pretrained(model_path)
        object.__setattr__(self, "_engine", None)

    def _mdhs_groups(self):
        return qkv_groups(self.bert)

    def _on_bind(self, store):
        object.__setattr__(self, "_engine", BertEngine(store, self.bert))

    def forward(self, input_ids, attention_mask=None):
        st = self.store(input_ids.device)
        B, S = input_ids.shape
        trainable = any(p.requires_grad for p in self.bert.encoder.parameters())
        need = trainable and torch.is_grad_enabled()
        h = _BertFn.apply(st.anchor, input_ids, attention_mask, self._engine, self.training, need, ())
        return h.view(B, S, h.shape[1])[:, 0, :]
