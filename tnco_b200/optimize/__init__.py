"""Core-object surface of the SA path (tnco/optimize/**): one optimizer object == one chain on the GPU."""
