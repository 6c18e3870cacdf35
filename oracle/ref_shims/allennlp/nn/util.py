"""allennlp 0.8.4 `nn.util` helpers used by the reference, restated from their published semantics.

Call sites in the reference: masked_mean (var_updown/var_updown/modules/updown_cell.py:6,266),
masked_softmax (updown-baseline/updown/modules/attention.py:6,93),
add_sentence_boundary_token_ids / sequence_cross_entropy_with_logits
(var_updown/var_updown/models/updown_captioner.py:12,265,464).
"""
import torch


def masked_softmax(vector, mask, dim=-1):
    # 0.8.4, memory_efficient=False branch: softmax(x*m)*m / (sum + 1e-13)
    if mask is None:
        return torch.softmax(vector, dim=dim)
    mask = mask.float()
    while mask.dim() < vector.dim():
        mask = mask.unsqueeze(1)
    result = torch.softmax(vector * mask, dim=dim)
    result = result * mask
    return result / (result.sum(dim=dim, keepdim=True) + 1e-13)


def masked_mean(vector, mask, dim, keepdim=False, eps=1e-8):
    one_minus_mask = ~mask.bool()
    replaced = vector.masked_fill(one_minus_mask, 0.0)
    value_sum = torch.sum(replaced, dim=dim, keepdim=keepdim)
    value_count = torch.sum(mask.float(), dim=dim, keepdim=keepdim)
    return value_sum / value_count.clamp(min=eps)


def add_sentence_boundary_token_ids(tensor, mask, sentence_begin_token, sentence_end_token):
    sequence_lengths = mask.sum(dim=1).detach().cpu().numpy()
    tensor_shape = list(tensor.data.shape)
    new_shape = list(tensor_shape)
    new_shape[1] = tensor_shape[1] + 2
    tensor_with_boundary_tokens = tensor.new_zeros(*new_shape)
    assert len(tensor_shape) == 2
    tensor_with_boundary_tokens[:, 1:-1] = tensor
    tensor_with_boundary_tokens[:, 0] = sentence_begin_token
    for i, j in enumerate(sequence_lengths):
        tensor_with_boundary_tokens[i, j + 1] = sentence_end_token
    new_mask = (tensor_with_boundary_tokens != 0).long()
    return tensor_with_boundary_tokens, new_mask


def sequence_cross_entropy_with_logits(logits, targets, weights, average="batch"):
    logits_flat = logits.view(-1, logits.size(-1))
    log_probs_flat = torch.log_softmax(logits_flat, dim=-1)
    targets_flat = targets.view(-1, 1).long()
    negative_log_likelihood_flat = -torch.gather(log_probs_flat, dim=1, index=targets_flat)
    negative_log_likelihood = negative_log_likelihood_flat.view(*targets.size())
    negative_log_likelihood = negative_log_likelihood * weights.float()
    if average == "batch":
        per_batch_loss = negative_log_likelihood.sum(1) / (weights.sum(1).float() + 1e-13)
        num_non_empty_sequences = (weights.sum(1) > 0).float().sum() + 1e-13
        return per_batch_loss.sum() / num_non_empty_sequences
    elif average == "token":
        return negative_log_likelihood.sum() / (weights.sum().float() + 1e-13)
    else:
        return negative_log_likelihood.sum(1) / (weights.sum(1).float() + 1e-13)


def min_value_of_dtype(dtype):
    return torch.finfo(dtype).min
