"""TEST INFRASTRUCTURE — stub of ``pettingzoo.ParallelEnv`` (a plain base class)."""


class ParallelEnv:
    metadata: dict = {}
    agents: list = []
    possible_agents: list = []
