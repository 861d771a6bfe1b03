"""CPU/torch restatement of the DSAN LMMD loss.  TEST INFRASTRUCTURE ONLY (see oracle/alignq_oracle.py header).

Follows cdf_alignment_admm/dsan_office/utils/mmd.py:9-41 (guassian_kernel, lmmd) and utils/Weight.py:4-59
(convert_to_onehot, Weight.cal_weight) line by line; pinned by oracle/make_mmd_golden.py, which imports the unmodified
reference in the build container, asserts bit-equality on CPU and writes tests/golden/mmd.npz."""
from __future__ import annotations

import numpy as np
import torch


def convert_to_onehot(sca_label, class_num=31):                      # Weight.py:4-5
    return np.eye(class_num)[sca_label]


def cal_weight(s_label, t_label, class_num=31):                      # Weight.py:10-59
    batch_size = s_label.size()[0]
    s_sca_label = s_label.cpu().data.numpy()
    s_vec_label = convert_to_onehot(s_sca_label, class_num)
    s_sum = np.sum(s_vec_label, axis=0).reshape(1, class_num)
    s_sum[s_sum == 0] = 100
    s_vec_label = s_vec_label / s_sum
    t_sca_label = t_label.cpu().data.max(1)[1].numpy()
    t_vec_label = t_label.cpu().data.numpy()
    t_sum = np.sum(t_vec_label, axis=0).reshape(1, class_num)
    t_sum[t_sum == 0] = 100
    t_vec_label = t_vec_label / t_sum
    weight_ss = np.zeros((batch_size, batch_size))
    weight_tt = np.zeros((batch_size, batch_size))
    weight_st = np.zeros((batch_size, batch_size))
    set_s, set_t = set(s_sca_label), set(t_sca_label)
    count = 0
    for i in range(class_num):
        if i in set_s and i in set_t:
            s_tvec = s_vec_label[:, i].reshape(batch_size, -1)
            t_tvec = t_vec_label[:, i].reshape(batch_size, -1)
            weight_ss = weight_ss + np.dot(s_tvec, s_tvec.T)
            weight_tt = weight_tt + np.dot(t_tvec, t_tvec.T)
            weight_st = weight_st + np.dot(s_tvec, t_tvec.T)
            count += 1
    if count != 0:
        weight_ss, weight_tt, weight_st = weight_ss / count, weight_tt / count, weight_st / count
    else:
        weight_ss = weight_tt = weight_st = np.array([0])
    return weight_ss.astype('float32'), weight_tt.astype('float32'), weight_st.astype('float32')


def guassian_kernel(source, target, kernel_mul=2.0, kernel_num=5, fix_sigma=None):      # mmd.py:9-22
    n_samples = int(source.size()[0]) + int(target.size()[0])
    total = torch.cat([source, target], dim=0)
    total0 = total.unsqueeze(0).expand(int(total.size(0)), int(total.size(0)), int(total.size(1)))
    total1 = total.unsqueeze(1).expand(int(total.size(0)), int(total.size(0)), int(total.size(1)))
    L2_distance = ((total0 - total1) ** 2).sum(2)
    if fix_sigma:
        bandwidth = fix_sigma
    else:
        bandwidth = torch.sum(L2_distance.data) / (n_samples ** 2 - n_samples)
    bandwidth /= kernel_mul ** (kernel_num // 2)
    bandwidth_list = [bandwidth * (kernel_mul ** i) for i in range(kernel_num)]
    kernel_val = [torch.exp(-L2_distance / bandwidth_temp) for bandwidth_temp in bandwidth_list]
    return sum(kernel_val)


def lmmd(source, target, s_label, t_label, kernel_mul=2.0, kernel_num=5, fix_sigma=None, class_num=31):   # mmd.py:24-41
    batch_size = source.size()[0]
    weight_ss, weight_tt, weight_st = cal_weight(s_label, t_label, class_num)
    dev, dt = source.device, source.dtype
    weight_ss = torch.from_numpy(weight_ss).to(dev).to(dt)
    weight_tt = torch.from_numpy(weight_tt).to(dev).to(dt)
    weight_st = torch.from_numpy(weight_st).to(dev).to(dt)
    kernels = guassian_kernel(source, target, kernel_mul=kernel_mul, kernel_num=kernel_num, fix_sigma=fix_sigma)
    loss = torch.zeros(1, dtype=dt, device=dev)
    if torch.sum(torch.isnan(sum(kernels))):
        return loss
    SS = kernels[:batch_size, :batch_size]
    TT = kernels[batch_size:, batch_size:]
    ST = kernels[:batch_size, batch_size:]
    loss = loss + torch.sum(weight_ss * SS + weight_tt * TT - 2 * weight_st * ST)
    return loss
