"""phyloforge_b200 — B200-native population-phylogeny hot path."""
