"""Same module layout as py-phylo2vec/phylo2vec/io: reader (load, read, load_newick), writer (save, write,
save_newick), plus load_batch / save_batch for device batches."""
from .reader import load, load_batch, load_newick, read
from .writer import save, save_batch, save_newick, write

__all__ = ["load", "read", "load_newick", "load_batch", "save", "write", "save_newick", "save_batch"]
