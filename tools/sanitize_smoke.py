"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): a few fused ticks in every mode
(stage1 auto-reset, stage2 group mode + scan FIFO, observe, raycast) and one learner step (TC and fp32 paths)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl_collision_avoidance_b200.model.net import Adam, CNNPolicy
from rl_collision_avoidance_b200.model.ppo import generate_train_data, ppo_update_stage1
from rl_collision_avoidance_b200.stage_world import StageWorld


def main():
    for scen, ar, beams in (('stage1', 1, 512), ('stage2', 2, 512), ('stage1', 1, 180)):
        env = StageWorld(beams, scenario=scen, num_worlds=2, seed=1, auto_reset=ar)
        env.reset_pose()
        st = [env.obs[:, None, :].repeat(1, 3, 1).contiguous(), torch.empty(env.N, 3, beams, device='cuda')]
        for t in range(6):
            env.control_vel(torch.rand(env.N, 2, device='cuda'), stack_in=st[t % 2], stack_out=st[(t + 1) % 2])
        env.raycast(env.state['pose'].clone())
        # the host-buffer call in its three traffic modes (kernel-written pinned host memory, mixed, DMA in world ranges)
        a_host = torch.rand(env.N, 2).pin_memory()
        for mode in (1, 2, 0):
            env.set_host_zero_copy(mode)
            env.step_host(a_host)
        torch.cuda.synchronize()
        env.close()
    if '--circle' in sys.argv:
        env = StageWorld(512, scenario='circle', num_worlds=1, seed=1, auto_reset=1)
        env.reset_pose()
        for t in range(2):
            env.control_vel(torch.rand(env.N, 2, device='cuda'))
        torch.cuda.synchronize()
        env.close()
    if '--env-only' in sys.argv:
        print('sanitize workload done (env only)')
        return
    for tc in (True, False):
        pol = CNNPolicy(max_batch=64)
        pol.set_tensor_cores(tc)
        opt = Adam(pol.parameters(), lr=5e-5)
        T, N = 4, 8
        obs = torch.rand(T, N, 3, 512, device='cuda') - 0.5
        mem = (obs, torch.rand(T, N, 2, device='cuda'), torch.rand(T, N, 2, device='cuda'), torch.rand(T, N, 2, device='cuda'),
               torch.rand(T, N, 1, device='cuda') - 1.0, None, torch.randn(T, N, device='cuda'), torch.randn(T, N, device='cuda'), None)
        tg, adv = generate_train_data(mem[7], 0.99, mem[6], torch.randn(N, device='cuda'), torch.rand(T, N, device='cuda') < 0.2, 0.95)
        mem = mem[:5] + (tg,) + mem[6:8] + (adv,)
        ppo_update_stage1(policy=pol, optimizer=opt, batch_size=32, memory=mem, epoch=1, coeff_entropy=5e-4, clip_value=0.1,
                          num_step=T, num_env=N, frames=3, obs_size=512, act_size=2)
        torch.cuda.synchronize()
    print('sanitize workload done')


if __name__ == '__main__':
    main()
