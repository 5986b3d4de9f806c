"""Oracle restatement of GLKS's vocabulary side (TEST INFRASTRUCTURE ONLY).

  Mixturer.forward ........... GLKS/Model.py:135-147   p = s * p_v + (1 - s) * (p_k @ onehot),  s = sigmoid(w . state + b)
  VocabGenerator tail ........ GLKS/Model.py:128-130   softmax(generator(feature))
"""
import torch
import torch.nn.functional as F


def mixture(sd, state, p_v, p_k, bmap, dense_onehot=False):
    """``sd``: {'linear1.weight' [1,H], 'linear1.bias' [1]}; ``bmap`` int64 [R,Lb]."""
    s = torch.sigmoid(F.linear(state.squeeze(1), sd['linear1.weight'], sd['linear1.bias']))
    if dense_onehot:
        oh = torch.zeros(bmap.size(0), bmap.size(1), p_v.size(1)).scatter_(2, bmap.unsqueeze(2), 1.0)
        copy = torch.bmm(p_k.unsqueeze(1), oh).squeeze(1)
    else:
        copy = torch.zeros_like(p_v).scatter_add_(1, bmap, p_k)
    return s * p_v + (1.0 - s) * copy


def vocab_head(weight, bias, feature):
    return torch.softmax(F.linear(feature, weight, bias), dim=-1)
