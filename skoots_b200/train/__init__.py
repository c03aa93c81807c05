"""Training-side pieces of SURVEY.md §8 row f4 (the augmentation step next to the hot path's training ops)."""
