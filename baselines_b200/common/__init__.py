"""Host-side pieces of the reference's `baselines.common` that the PPO2 / DQN hot path touches."""
