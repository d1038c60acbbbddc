"""ResNet_34 embedding + residual knowledge distillation (model/resnet.py, distill_main.py:59-74 of the reference):
the native network program behind the drop-in module against the CPU oracle on identical weights and inputs.

Tolerances: the embedding / stage features / losses / gradients are produced with bf16 storage; as for FSRNet the
oracle's own bf16-storage evaluation (``precision="bf16"``) is the yardstick for the deviation that storage rounding
makes unavoidable through 36 train-mode BatchNorms, with the 1e-2 bf16 budget of BASELINE.json on top."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu

SLACK = 1.6
B = 8


def _nets():
    from crfr_b200.model.resnet import ResNet_34
    from oracle import resnet_oracle as RO
    nets, sds = [], []
    for i, seed in enumerate((77, 78, 79)):            # teacher, student, assistant (as in oracle/make_golden.py)
        torch.manual_seed(seed)
        net = ResNet_34()
        sd = RO.randomize_norm_params(RO.build_resnet34_state_dict(seed), 100 + i)
        net.load_state_dict(sd)
        nets.append(net.cuda())
        sds.append(sd)
    return nets, sds


def test_forward_train_eval_and_running_stats(cuda, golden_dir):
    from oracle import resnet_oracle as RO
    (teacher, student, _), sds = _nets()
    g = np.load(golden_dir + "/resnet34.npz")
    x = RO.synthetic_faces(B)
    # eval mode (teacher): running statistics
    teacher.eval()
    with torch.no_grad():
        t_outs = teacher(x.cuda())
    ref = RO.resnet34_forward(sds[0], x, training=False)
    emu = RO.resnet34_forward(sds[0], x, training=False, pr=RO.Precision("bf16"))
    assert rel_err(ref[0], torch.from_numpy(g["emb_t"])) < 1e-3              # oracle == reference fixture
    for o, r, e in zip(t_outs, ref, emu):
        assert o.dtype == torch.float32 and o.shape == r.shape and torch.isfinite(o).all()
        assert rel_err(o, r) < SLACK * rel_err(e, r) + 1e-2, (rel_err(o, r), rel_err(e, r))
    # train mode (student): batch statistics + buffer update
    student.train()
    nb = {}
    ref = RO.resnet34_forward(sds[1], x, training=True, new_buffers=nb)
    emu = RO.resnet34_forward(sds[1], x, training=True, pr=RO.Precision("bf16"))
    s_outs = student(x.cuda())
    for o, r, e in zip(s_outs, ref, emu):
        assert rel_err(o, r) < SLACK * rel_err(e, r) + 1e-2, (rel_err(o, r), rel_err(e, r))
    new_sd = student.state_dict()
    for k in ("bn1.running_mean", "bn1.running_var", "layer2.0.downsample.1.running_var", "layer4.2.bn2.running_mean",
              "bn_o1.running_var", "bn_o2.running_mean", "bn_o2.running_var"):
        # the BatchNorm1d statistics are means of 8 bf16-rounded fc outputs: looser than the image-sized BatchNorms
        assert rel_err(new_sd[k], nb[k]) < (6e-2 if k.startswith("bn_o2") else 2e-2), k
    assert int(new_sd["bn1.num_batches_tracked"]) == 1 and int(new_sd["layer3.5.bn2.num_batches_tracked"]) == 1
    assert rel_err(new_sd["bn1.running_mean"], torch.from_numpy(g["bn1_running_mean"])) < 2e-2


def test_kd_step_losses_and_gradients(cuda, golden_dir):
    """distill_main.py:59-74 on one forward: L_s, L_a, dL_s/dtheta_S, dL_a/dtheta_A and dL_a/dtheta_S."""
    from crfr_b200.loss import MSELoss, ResidualKDLoss
    from oracle import resnet_oracle as RO
    (teacher, student, assistant), sds = _nets()
    g = np.load(golden_dir + "/resnet34.npz")
    x = RO.synthetic_faces(B)
    teacher.eval(); student.train(); assistant.train()
    xc = x.cuda()
    with torch.no_grad():
        t_outs = teacher(xc)
    s_outs = student(xc)
    a_outs = assistant(xc)
    mse, kd = MSELoss(), ResidualKDLoss()
    l_s = mse(s_outs[0], t_outs[0].detach())
    l_a = sum(kd(t_outs[k], s_outs[k], a_outs[k]) for k in (1, 2, 3, 4)) + kd(t_outs[0], s_outs[0], a_outs[0])
    names = [k for k, _ in student.named_parameters()]
    g_s = torch.autograd.grad(l_s, list(student.parameters()), retain_graph=True)
    g_a = torch.autograd.grad(l_a, list(assistant.parameters()), retain_graph=True)
    g_as = torch.autograd.grad(l_a, list(student.parameters()))
    torch.cuda.synchronize()

    r_ls, r_la, rg_s, rg_a, rg_as, _ = RO.kd_step(sds[0], sds[1], sds[2], x)
    e_ls, e_la, eg_s, eg_a, eg_as, _ = RO.kd_step(sds[0], sds[1], sds[2], x, precision="bf16")
    assert abs(r_ls.item() - float(g["l_s"])) < 1e-4 * abs(float(g["l_s"]))   # oracle == reference fixture
    assert abs(r_la.item() - float(g["l_a"])) < 1e-4 * abs(float(g["l_a"]))
    for ours, r, e in ((l_s, r_ls, e_ls), (l_a, r_la, e_la)):
        assert abs(ours.item() - r.item()) < SLACK * abs(e.item() - r.item()) + 1e-2 * abs(r.item())
    for tag, ours_g, ref_g, emu_g in (("s", g_s, rg_s, eg_s), ("a", g_a, rg_a, eg_a), ("as", g_as, rg_as, eg_as)):
        flat_o, flat_r, flat_e = [], [], []
        for k, og in zip(names, ours_g):
            assert torch.isfinite(og).all(), (tag, k)
            if k in RO.RESNET_NULL_GRAD:
                continue
            e_ours, e_emu = rel_err(og, ref_g[k]), rel_err(emu_g[k], ref_g[k])
            assert e_ours < SLACK * e_emu + 2e-2, (tag, k, e_ours, e_emu)
            flat_o.append(og.flatten().cpu().double()); flat_r.append(ref_g[k].flatten().double())
            flat_e.append(emu_g[k].flatten().double())
        a, b, c = torch.cat(flat_o), torch.cat(flat_r), torch.cat(flat_e)
        cos_ours, cos_emu = float(a @ b / (a.norm() * b.norm())), float(c @ b / (c.norm() * b.norm()))
        # direction and length of the whole gradient: as good as the bf16-storage evaluation of the oracle itself
        # (with 8 samples behind the BatchNorm1d that evaluation is ~0.8 from fp32 on the student loss)
        assert cos_ours > min(0.98, cos_emu - 0.02), (tag, cos_ours, cos_emu)
        assert abs(a.norm().item() - b.norm().item()) < max(5e-2 * b.norm().item(),
                                                            SLACK * abs(c.norm().item() - b.norm().item())), tag


def test_native_kd_step_matches_module_path(cuda):
    """crfr_kd_train_step == the same step composed from the drop-in modules and autograd (same kernels; the native
    call only skips the fp32 NCHW round trip of the outputs and seeds the feature gradients in bf16 directly)."""
    from crfr_b200.loss import MSELoss, ResidualKDLoss
    from crfr_b200.model.resnet import kd_train_step
    from oracle import resnet_oracle as RO
    (teacher, student, assistant), sds = _nets()
    teacher.eval(); student.train(); assistant.train()
    xc = RO.synthetic_faces(B).cuda()
    with torch.no_grad():
        t_outs = teacher(xc)
    s_outs, a_outs = student(xc), assistant(xc)
    mse, kd = MSELoss(), ResidualKDLoss()
    l_s = mse(s_outs[0], t_outs[0])
    l_a = sum(kd(t_outs[k], s_outs[k], a_outs[k]) for k in (1, 2, 3, 4)) + kd(t_outs[0], s_outs[0], a_outs[0])
    (l_s + l_a).backward()
    ref_s = [p.grad.clone() for p in student.parameters()]
    ref_a = [p.grad.clone() for p in assistant.parameters()]
    for net, sd in zip((student, assistant), sds[1:]):      # undo the BatchNorm buffer update of the first forward
        net.load_state_dict(sd)
        net.zero_grad(set_to_none=True)
    losses = kd_train_step(teacher, student, assistant, xc)
    torch.cuda.synchronize()
    assert abs(losses[0].item() - l_s.item()) < 1e-4 * abs(l_s.item())
    assert abs(losses[1].item() - l_a.item()) < 1e-4 * abs(l_a.item())
    names = [k for k, _ in student.named_parameters()]
    # the two paths round the summed embedding gradient differently (autograd sums in fp32 and rounds once, the native
    # step sums two bf16 slots); 36 BatchNorm backward passes amplify that to a few per cent on the deepest tensors
    for tag, net, ref in (("s", student, ref_s), ("a", assistant, ref_a)):
        fo, fr = [], []
        for k, p, r in zip(names, net.parameters(), ref):
            if k in RO.RESNET_NULL_GRAD:
                continue
            assert rel_err(p.grad, r) < 6e-2, (tag, k, rel_err(p.grad, r))
            fo.append(p.grad.flatten().double()); fr.append(r.flatten().double())
        a, b = torch.cat(fo), torch.cat(fr)
        assert float(a @ b / (a.norm() * b.norm())) > 0.999, tag
    assert int(student.state_dict()["bn1.num_batches_tracked"]) == 1


def test_ir50_teacher_eval_forward(cuda, golden_dir):
    """The frozen IR_50 teacher (DISTILLATION/model/model_irse.py, distill_main.py:201) in eval mode."""
    from crfr_b200.model.model_irse import IR_50
    from oracle import resnet_oracle as RO
    torch.manual_seed(91)
    net = IR_50([112, 112])
    sd = RO.randomize_bn_everywhere(RO.build_ir50_state_dict(91), 191)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = RO.synthetic_faces(4, seed=4322)
    with torch.no_grad():
        emb = net(x.cuda())
        ref = RO.ir50_forward(sd, x)
        emu = RO.ir50_forward(sd, x, pr=RO.Precision("bf16"))
    g = np.load(golden_dir + "/ir50.npz")
    assert rel_err(ref, torch.from_numpy(g["emb"])) < 1e-4                    # oracle == reference fixture
    assert emb.shape == (4, 512) and torch.isfinite(emb).all()
    assert rel_err(emb, ref) < SLACK * rel_err(emu, ref) + 1e-2, (rel_err(emb, ref), rel_err(emu, ref))
    net.train()
    with pytest.raises(RuntimeError):
        net(x.cuda())


def test_rejects_unsupported(cuda):
    from crfr_b200.model.resnet import ResNet_34
    net = ResNet_34().cuda()
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 3, 112, 112))
    with pytest.raises(ValueError):
        net(torch.zeros(2, 3, 96, 96, device="cuda"))


def test_ir50_stage_features(cuda):
    """IR_50 as a feature extractor (DISTILLATION/model/utils.py:36-52): the four body-stage outputs next to the embedding;
    they have the shapes of ResNet_34's x1..x4, which is what distill_main.py:59 unpacks from its teacher."""
    from crfr_b200.model.model_irse import IR_50
    from oracle import resnet_oracle as RO
    torch.manual_seed(91)
    net = IR_50([112, 112])
    sd = RO.randomize_bn_everywhere(RO.build_ir50_state_dict(91), 191)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = RO.synthetic_faces(2, seed=4322)
    with torch.no_grad():
        outs = net.forward_features(x.cuda())
        emb = net(x.cuda())
        ref_emb, ref = RO.ir50_forward(sd, x, want_features=True)
        _, emu = RO.ir50_forward(sd, x, pr=RO.Precision("bf16"), want_features=True)
    assert torch.equal(outs[0], emb)
    for f, r, e, shape in zip(outs[1:], ref, emu, ((64, 56), (128, 28), (256, 14), (512, 7))):
        assert tuple(f.shape) == (2, shape[0], shape[1], shape[1])
        assert rel_err(f, r) < SLACK * rel_err(e, r) + 1e-2, (shape, rel_err(f, r), rel_err(e, r))


@pytest.mark.parametrize("teacher_kind", ["resnet34", "ir50"])
def test_kd_step_hr_teacher_lr_student(cuda, teacher_kind):
    """'HR teacher / LR student' (BASELINE configs[2]): the teacher sees x_hr, the student and the assistant x_lr, with a
    ResNet_34 or an IR_50 teacher (whose stage outputs are the t_k): the native step == the same step composed from the
    drop-in modules, the loss modules and autograd."""
    from crfr_b200.loss import MSELoss, ResidualKDLoss
    from crfr_b200.model.model_irse import IR_50
    from crfr_b200.model.resnet import kd_train_step
    from oracle import resnet_oracle as RO
    (teacher, student, assistant), sds = _nets()
    if teacher_kind == "ir50":
        torch.manual_seed(91)
        teacher = IR_50([112, 112])
        teacher.load_state_dict(RO.randomize_bn_everywhere(RO.build_ir50_state_dict(91), 191))
        teacher = teacher.cuda()
    teacher.eval(); student.train(); assistant.train()
    x_hr = RO.synthetic_faces(B).cuda()
    x_lr = RO.synthetic_faces(B, seed=999).cuda()
    with torch.no_grad():
        t_outs = teacher.forward_features(x_hr) if teacher_kind == "ir50" else teacher(x_hr)
    s_outs, a_outs = student(x_lr), assistant(x_lr)
    mse, kd = MSELoss(), ResidualKDLoss()
    l_s = mse(s_outs[0], t_outs[0])
    l_a = sum(kd(t_outs[k], s_outs[k], a_outs[k]) for k in (1, 2, 3, 4)) + kd(t_outs[0], s_outs[0], a_outs[0])
    (l_s + l_a).backward()
    ref_s = [p.grad.clone() for p in student.parameters()]
    ref_a = [p.grad.clone() for p in assistant.parameters()]
    for net, sd in zip((student, assistant), sds[1:]):
        net.load_state_dict(sd)
        net.zero_grad(set_to_none=True)
    losses = kd_train_step(teacher, student, assistant, x_hr, x_lr=x_lr)
    torch.cuda.synchronize()
    assert abs(losses[0].item() - l_s.item()) < 1e-4 * abs(l_s.item())
    assert abs(losses[1].item() - l_a.item()) < 1e-4 * abs(l_a.item())
    # and the LR input really is what the student saw: the same call on x_hr alone gives different losses
    for net, sd in zip((student, assistant), sds[1:]):
        net.load_state_dict(sd)
    same = kd_train_step(teacher, student, assistant, x_hr)
    assert same[0].item() != losses[0].item() and same[1].item() != losses[1].item()
    names = [k for k, _ in student.named_parameters()]
    for net, sd in zip((student, assistant), sds[1:]):
        net.load_state_dict(sd)
    kd_train_step(teacher, student, assistant, x_hr, x_lr=x_lr)
    for tag, net, ref in (("s", student, ref_s), ("a", assistant, ref_a)):
        fo, fr = [], []
        for k, p, r in zip(names, net.parameters(), ref):
            if k in RO.RESNET_NULL_GRAD:
                continue
            assert rel_err(p.grad, r) < 6e-2, (tag, k, rel_err(p.grad, r))
            fo.append(p.grad.flatten().double()); fr.append(r.flatten().double())
        a, b = torch.cat(fo), torch.cat(fr)
        assert float(a @ b / (a.norm() * b.norm())) > 0.999, tag


@pytest.mark.parametrize("use_graph", [False, True])
def test_kd_trainer_step_equals_native_step_plus_rmsprop(cuda, use_graph):
    """KDTrainer (flat arenas, fused RMSprop, the data-parallel plumbing with world = 1): its gradients are those of
    kd_train_step, and its update is torch.optim.RMSprop's with the reference's hyper-parameters (distill_main.py:222-225)
    applied to exactly those gradients.  (The first RMSprop step moves every element by ~10 lr whatever its size, so
    parameters are compared through the update formula on the trainer's own gradients, not across two runs whose
    near-zero gradient elements may differ in sign.)"""
    from crfr_b200.model.resnet import kd_train_step
    from crfr_b200.trainer import KDTrainer
    from oracle import resnet_oracle as RO
    x_hr, x_lr = RO.synthetic_faces(B).cuda(), RO.synthetic_faces(B, seed=5).cuda()
    (teacher, student, assistant), sds = _nets()
    teacher.eval(); student.train(); assistant.train()
    ref_losses = kd_train_step(teacher, student, assistant, x_hr, x_lr=x_lr).clone()
    ref_g = [[p.grad.clone() for p in n.parameters()] for n in (student, assistant)]
    (teacher2, student2, assistant2), _ = _nets()
    teacher2.eval(); student2.train(); assistant2.train()
    tr = KDTrainer(teacher2, student2, assistant2, lr=1e-4, use_graph=use_graph)
    before = [f.flat_p.clone() for f in (tr.S, tr.A)]
    losses = tr.step(x_hr, x_lr)
    torch.cuda.synchronize()
    assert torch.allclose(losses, ref_losses, rtol=1e-5)
    for f, r, p0 in zip((tr.S, tr.A), ref_g, before):
        for gv, g in zip(f.grad_views, r):
            assert rel_err(gv, g) < 1e-3
        g = f.flat_g + 1e-5 * p0                                   # weight decay
        sq = 0.01 * g * g                                          # (1 - alpha) g^2 from a zero state
        expect = p0 - 1e-4 * g / (sq.sqrt() + 1e-8)
        assert torch.allclose(f.flat_p, expect, rtol=1e-5, atol=1e-7)
        assert torch.allclose(f.flat_sq, sq, rtol=1e-5, atol=1e-12)
    # parameters are views of the arena: the modules see the update
    assert student2.conv1.weight.data_ptr() == tr.S.flat_p.data_ptr()
    assert int(student2.state_dict()["bn1.num_batches_tracked"]) == 1
