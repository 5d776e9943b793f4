"""Drop-in package: same import paths as neuronest/custom-alphazero for the self-play hot path."""
