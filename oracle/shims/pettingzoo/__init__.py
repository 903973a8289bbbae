class ParallelEnv:
    possible_agents = []  # survives the reference's ctor-order bug
    agents = []
