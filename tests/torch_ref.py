"""Plain PyTorch fp32 reference of the policy network and PPO loss (test-only; independent of the product).
Architecture from SURVEY.md App. B / model/net.py:16-34."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefPolicy(nn.Module):
    def __init__(self):
        super().__init__()
        self.logstd = nn.Parameter(torch.zeros(2))
        for p in ('act', 'crt'):
            setattr(self, p + '_fea_cv1', nn.Conv1d(3, 32, 5, 2, 1))
            setattr(self, p + '_fea_cv2', nn.Conv1d(32, 32, 3, 2, 1))
            setattr(self, p + '_fc1', nn.Linear(4096, 256))
            setattr(self, p + '_fc2', nn.Linear(260, 128))
        self.actor1 = nn.Linear(128, 1)
        self.actor2 = nn.Linear(128, 1)
        self.critic = nn.Linear(128, 1)

    def tower(self, p, x, goal, speed):
        h = F.relu(getattr(self, p + '_fea_cv1')(x))
        h = F.relu(getattr(self, p + '_fea_cv2')(h))
        h = F.relu(getattr(self, p + '_fc1')(h.flatten(1)))
        return F.relu(getattr(self, p + '_fc2')(torch.cat((h, goal, speed), -1)))

    def forward(self, x, goal, speed):
        a = self.tower('act', x, goal, speed)
        mean = torch.cat((torch.sigmoid(self.actor1(a)), torch.tanh(self.actor2(a))), -1)
        v = self.critic(self.tower('crt', x, goal, speed))
        return v, mean

    def logprob(self, mean, action):
        var = torch.exp(2 * self.logstd)
        return (-(action - mean) ** 2 / (2 * var) - 0.5 * math.log(2 * math.pi) - self.logstd).sum(-1, keepdim=True)


def ppo_loss(pol, x, goal, speed, action, old_lp, adv, target, clip=0.1, coeff=5e-4, vcoef=20.0):
    v, mean = pol(x, goal, speed)
    lp = pol.logprob(mean, action)
    ratio = torch.exp(lp - old_lp.view(-1, 1))
    A = adv.view(-1, 1)
    pl = -torch.min(ratio * A, torch.clamp(ratio, 1 - clip, 1 + clip) * A).mean()
    vl = F.mse_loss(v, target.view(-1, 1))
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + pol.logstd).sum()
    return pl + vcoef * vl - coeff * ent, (pl, vl, ent), v, mean
