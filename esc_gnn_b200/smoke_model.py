"""One tiny forward + backward of the ZINC NestedGIN_eff on cuda:0 (called by __graft_entry__.smoke())."""
import torch


def run():
    import numpy as np
    from . import synth, zinc_model
    from .batch import Batch
    from .data import Data
    from .transform import encode_batch_host
    graphs = []
    for i in range(8):
        g = synth.make_graph(2, i)
        r = encode_batch_host(g['edge_index'][0], g['edge_index'][1], np.array([0, g['edge_index'].shape[1]]),
                              np.array([0, g['num_nodes']]), 3, True, False, local_ordinals=True)
        graphs.append(Data(x=torch.as_tensor(g['x']), edge_index=r.edge_index, edge_attr=torch.as_tensor(g['edge_attr']),
                           y=torch.as_tensor(g['y']), pos_enc=r.pos_enc, pos_index=r.pos_index, pos_batch=r.pos_batch))
    batch = Batch.from_data_list(graphs).to('cuda')
    torch.manual_seed(0)
    model = zinc_model.NestedGIN_eff(None, 3, hidden=64).cuda()
    loss = torch.nn.L1Loss()(model(batch), batch.y.view(-1, 1))
    loss.backward()
    assert torch.isfinite(loss) and model.z_initial.weight.grad.abs().sum() > 0
    return loss.item()
