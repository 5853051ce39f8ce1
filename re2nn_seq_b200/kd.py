"""Teacher-mixing losses used when marryup_type is 'kd' or 'pr' (src_seq/baselines/KD.py:3-18).
Out of the hot-path scope: plain torch on the op's all_scores output."""
import torch
import torch.nn.functional as F


def KD_loss(scores, re_scores, args):
    temp = args.c1_kdpr
    student = torch.log_softmax(scores / temp, 2)
    teacher = torch.softmax(re_scores / temp, 2)
    return F.kl_div(student, teacher, reduction='mean') * temp * temp


def PR_loss(scores, re_scores, args):
    log_student = torch.log_softmax(scores, 2)
    student = torch.softmax(scores, 2)
    teacher = torch.softmax(student * (torch.exp(re_scores - 1) * args.c1_kdpr), 2)
    return F.kl_div(log_student, teacher, reduction='mean')


def ml_loss(all_scores, lengths, labels, margin):
    """local_loss_func == 'ML': nn.MultiMarginLoss(margin) over the valid positions (model_decompose.py:84-85,
    model_onehot.py:61-62 with the flatten of utils.py:153-164); a device-side reduction over N x C scores."""
    B, L, _ = all_scores.shape
    mask = torch.arange(L, device=all_scores.device).unsqueeze(0) < lengths.unsqueeze(1)
    return F.multi_margin_loss(all_scores[mask], labels[:, :L][mask], margin=float(margin))
