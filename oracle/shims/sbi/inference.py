import torch


def simulate_for_sbi(simulator, proposal, num_simulations, simulation_batch_size=None, **_kw):
    theta = proposal.sample((num_simulations,))
    bs = simulation_batch_size or num_simulations
    xs = [simulator(theta[i:i + bs]) for i in range(0, num_simulations, bs)]
    return theta, torch.cat(xs, 0)
