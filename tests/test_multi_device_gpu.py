"""A process that touches a second GPU (per-device shared-memory opt-in; round 2: the cheap device context `ops._on`): a model
on cuda:1 while cuda:0 is the current device gives the bits of the same model on cuda:0, and the current device is restored.  Skipped on a
one-GPU box."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def test_forward_on_a_non_current_device():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from deeprecommendation_b200.neural_collaborative_filtering.models import AttentionNCF, BasicNCF
    torch.cuda.set_device(0)
    kw = dict(item_dim=2094, user_dim=2094, item_emb=128, user_emb=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    sd = synth.to_torch(synth.basic_ncf_weights(seed=1, **kw))
    xi = torch.from_numpy(synth.item_profiles(300, seed=2))
    xu = torch.from_numpy((synth.item_profiles(300, seed=3) - 0.3) * 0.1)
    outs = []
    for dev in ('cuda:0', 'cuda:1'):
        m = BasicNCF(**kw).to(dev).eval()
        m.load_state_dict(sd)
        with torch.no_grad():
            outs.append(m(xu.to(dev), xi.to(dev)).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])
    kw = dict(item_dim=512, item_emb=128, user_emb=128, att_dense=128, mlp_dense_layers=[256, 128], dropout_rate=0.2)
    sd = synth.to_torch(synth.attention_ncf_weights(seed=4, **kw))
    prof = torch.from_numpy(synth.item_profiles(700, seed=5, f_binary=256, f_dense=256))
    um = ((torch.rand(64, 600) < 0.2) * (torch.randint(1, 11, (64, 600)) * 0.5 - 2.75)).float()
    outs = []
    for dev in ('cuda:0', 'cuda:1'):
        m = AttentionNCF(**kw).to(dev).eval()
        m.load_state_dict(sd)
        with torch.no_grad():
            outs.append(m(prof[600:664].to(dev), prof[:600].to(dev), um.to(dev)).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])
