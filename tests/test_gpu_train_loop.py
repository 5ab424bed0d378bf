"""The reference's training loop, /root/reference/train.py:133-150, run against the mirror: iterate -> set_progress ->
train_step -> metric on the returned preds, for one epoch of four batches, compared step by step with the same loop replayed on the
CPU oracle (numpy legacy stream for the shuffle and the samples, oracle forward, torch autograd + clip_grad_norm 5 + Adam at the
scheduled learning rate).  GPU only.

Bars: seed batches and sampled ids bit-exact; metric within 1e-6 of sklearn on the oracle's predictions unless a prediction flips
(checked through the logits, rtol 1e-3); parameters after every step rtol 2e-3 / atol 2e-5 on >= 99.9 % of
the elements (Adam's update of a weight whose gradient is ~eps is rounding noise; those stay within one learning-rate step)."""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from oracle import layers, metrics as ometrics, sampler as osampler
from oracle.mt19937 import MT19937Oracle

pytestmark = pytest.mark.gpu


def _problem(g):
    from pytorch_graphsage_b200 import synth
    from pytorch_graphsage_b200.problem import NodeProblem
    prob = synth.make_problem('tiny', seed=4)
    n = prob['n_nodes']
    folds = np.array(['val'] * n, dtype=object)
    folds[1:601] = 'train'
    folds[0] = 'dummy'
    problem = NodeProblem(task='classification', n_classes=prob['n_classes'], feats=prob['feats'], folds=folds.astype(str),
                          targets=prob['targets'], adj=synth.triplets(prob['adj']), train_adj=synth.triplets(prob['adj']), sparse=True)
    return prob, problem


def test_reference_training_loop_runs_verbatim_and_matches_the_oracle():
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200 import aggregator_lookup, prep_lookup, sampler_lookup, GSSupervised
    from pytorch_graphsage_b200.helpers import set_seeds, to_numpy
    seed, epochs, batch_size = 123, 1, 200
    set_seeds(seed)                                                             # train.py:81
    prob, problem = _problem(g)
    model = GSSupervised(**{                                                    # train.py:94-123
        "sampler_class": sampler_lookup['sparse_uniform_neighbor_sampler'], "adj": problem.adj, "train_adj": problem.train_adj,
        "prep_class": prep_lookup['identity'], "aggregator_class": aggregator_lookup['mean'],
        "input_dim": problem.feats_dim, "n_nodes": problem.n_nodes, "n_classes": problem.n_classes,
        "layer_specs": [
            {"n_train_samples": 25, "n_val_samples": 25, "output_dim": 128, "activation": F.relu},
            {"n_train_samples": 10, "n_val_samples": 10, "output_dim": 128, "activation": lambda x: x},
        ],
        "lr_init": 0.01, "lr_schedule": 'linear', "weight_decay": 0.0,
    })
    model = model.cuda()
    ref = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ropt = torch.optim.Adam(list(ref.values()), lr=0.01)

    # ---- the oracle's replay of the same loop --------------------------------------------------------------------
    adj = prob['adj']
    indptr, data, shape = adj['indptr'], adj['data'], adj['shape']
    indices = np.arange(data.shape[0]) - np.repeat(indptr[:-1], np.diff(indptr))
    deg = osampler.row_degrees(indptr, data)
    rs = MT19937Oracle(seed ** 2)
    nodes = np.where(problem.folds == 'train')[0]
    idx = rs.permutation(nodes.shape[0])                                       # problem.py:146
    chunks = np.array_split(idx, idx.shape[0] // batch_size + 1)
    feats_cpu, targets_cpu = torch.from_numpy(prob['feats']), torch.from_numpy(prob['targets'].reshape(-1))

    # ---- train.py:133-150, verbatim but for `problem.metric_fn(targets, preds)` taking the device tensors ----------
    set_seeds(seed ** 2)
    steps = 0
    for epoch in range(epochs):
        _ = model.train()
        for ids, targets, epoch_progress in problem.iterate(mode='train', shuffle=True, batch_size=batch_size):
            model.set_progress((epoch + epoch_progress) / epochs)
            preds = model.train_step(
                ids=ids,
                feats=problem.feats,
                targets=targets,
                loss_fn=problem.loss_fn,
            )
            train_metric = problem.metric_fn(targets, preds)
            host_metric = problem.metric_fn(to_numpy(targets), to_numpy(preds))  # train.py:150 as written: numpy in, same numbers out

            # the oracle's step on the same batch
            want_ids = nodes[chunks[steps]]
            assert np.array_equal(ids.cpu().numpy(), want_ids), 'seed batch %d differs from np.random.permutation + array_split' % steps
            ids1 = osampler.sparse_sample(indptr, indices, data, shape, deg, want_ids, 25, rs.randint)
            ids2 = osampler.sparse_sample(indptr, indices, data, shape, deg, ids1, 10, rs.randint)
            assert np.array_equal(model.peek('ids2').cpu().numpy(), ids2), 'sampled ids differ at step %d' % steps
            lr = 0.01 * float(1 - (epoch + epoch_progress) / epochs) / 1               # LRSchedule.linear, epochs = 1 (models.py:66)
            assert abs(model.lr - lr) < 1e-15
            for grp in ropt.param_groups:
                grp['lr'] = lr
            ropt.zero_grad()
            logits = layers.forward_stack([torch.from_numpy(a) for a in (want_ids, ids1, ids2)], feats_cpu, ref)
            np.testing.assert_allclose(preds.cpu().numpy(), logits.detach().numpy(), rtol=1e-3, atol=1e-4)
            loss = F.cross_entropy(logits, targets_cpu[want_ids])
            assert abs(model.last_loss.item() - loss.item()) < 1e-3
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(ref.values()), 5)
            ropt.step()
            for name, p in model.named_parameters():
                # Adam moves a weight by lr * g / (sqrt(v) + eps): where |g| ~ eps the step is decided by rounding noise, so a
                # handful of elements may land anywhere within one step (lr) of the oracle's; everything else is tight
                got, want = p.detach().cpu().numpy(), ref[name].detach().numpy()
                off = np.abs(got - want) > 2e-5 + 2e-3 * np.abs(want)
                assert off.mean() <= 1e-3, '%s after step %d: %.4f %% of the elements differ' % (name, steps, 100 * off.mean())
                np.testing.assert_allclose(got, want, rtol=0, atol=(steps + 1) * 0.011, err_msg='%s after step %d' % (name, steps))
            want_metric = ometrics.classification(targets_cpu[want_ids].numpy().reshape(-1, 1), preds.cpu().numpy())
            assert train_metric == host_metric
            assert abs(train_metric['micro'] - want_metric['micro']) < 1e-9 and abs(train_metric['macro'] - want_metric['macro']) < 1e-9
            steps += 1
    assert steps == 4
    model.check()


def test_evaluate_loop_matches_the_oracle():
    """train.py:29-36 `evaluate`: iterate(shuffle=False) -> model(ids, feats, train=False) -> metric over the stacked predictions."""
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200 import aggregator_lookup, prep_lookup, sampler_lookup, GSSupervised
    from pytorch_graphsage_b200.helpers import set_seeds
    set_seeds(5)
    prob, problem = _problem(g)
    model = GSSupervised(input_dim=problem.feats_dim, n_nodes=problem.n_nodes, n_classes=problem.n_classes,
                         layer_specs=[dict(n_train_samples=25, n_val_samples=25, output_dim=128, activation=F.relu),
                                      dict(n_train_samples=10, n_val_samples=10, output_dim=128, activation=lambda x: x)],
                         aggregator_class=aggregator_lookup['mean'], prep_class=prep_lookup['identity'],
                         sampler_class=sampler_lookup['sparse_uniform_neighbor_sampler'], adj=problem.adj, train_adj=problem.train_adj).cuda()
    _ = model.eval()
    set_seeds(25)
    preds, acts = [], []
    for (ids, targets, _) in problem.iterate(mode='val', shuffle=False):
        preds.append(model(ids, problem.feats, train=False))
        acts.append(targets)
    got = problem.metric_fn(torch.cat(acts), torch.cat(preds))
    want = ometrics.classification(torch.cat(acts).cpu().numpy(), torch.cat(preds).cpu().numpy())
    assert abs(got['micro'] - want['micro']) < 1e-9 and abs(got['macro'] - want['macro']) < 1e-9
