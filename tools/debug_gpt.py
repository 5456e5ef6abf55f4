import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from oracle import models as M
nt.init(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nt.set_gemm_mode(mode)
cfg = M.GPTConfig(batch=2, context=128, embed_dim=128, heads=2, blocks=2, vocab=50)
params = M.gpt_init(cfg, 0)
tokens, labels = M.synthetic_gpt_batch(cfg, 1)
ref = M.gpt_step(params, None, tokens, labels, 0.05, cfg, adam=False)
layers = trainer.build_gpt_layers(2, 128, 128, 2, 2, 50, adam=False, seed=0)
nt.install()
from net.loss import cross_entropy_loss
from gpt.attdecoderblock import attdecoderblock_layer
x = tokens
for l in layers:
    x = l.forward(x, M.causal_mask(128)) if isinstance(l, attdecoderblock_layer) else l.forward(x)
flat = np.reshape(x, (-1, 50))
oh = np.zeros(flat.shape); oh[np.arange(flat.shape[0]), labels.reshape(-1)] = 1
loss, delta, pred = cross_entropy_loss(flat, oh)
print("loss", float(loss), ref["loss"])
delta = np.reshape(delta, x.shape)
for l in reversed(layers):
    delta = l.backward(delta)
grads = [layers[0].text_embedding.delta, layers[0].pos_embedding.delta]
for blk in layers[1:3]:
    for c in blk._children():
        grads += [c.gamma_delta, c.beta_delta] if hasattr(c, "gamma") else [c.params_delta, c.bias_delta]
grads += [layers[3].gamma_delta, layers[3].beta_delta, layers[4].fc0.params_delta, layers[4].fc0.bias_delta, layers[4].fc1.params_delta, layers[4].fc1.bias_delta]
for i, (a, b) in enumerate(zip(grads, M.tree_leaves(ref["grads"]))):
    a = np.asarray(a).reshape(-1); b = b.reshape(-1)
    e = np.max(np.abs(a - b)) / np.max(np.abs(b))
    print(i, b.shape, "%.2e" % e, "max|g| %.2e" % np.max(np.abs(b)))
