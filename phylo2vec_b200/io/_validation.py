"""File validation of the I/O functions; mirrors py-phylo2vec/phylo2vec/io/_validation.py."""

from pathlib import Path

FILE_EXTENSIONS = {
    "array": [".csv", ".txt"],
    "newick": [".txt", ".nwk", ".newick", ".tree", ".treefile"],
}


def check_path(filepath, file_type: str) -> None:
    suffix = Path(filepath).suffix
    file_extensions = FILE_EXTENSIONS[file_type]
    if suffix not in file_extensions:
        raise ValueError(
            f"Unsupported file extension: {suffix}. "
            f"Accepted extensions for {file_type} files: {file_extensions}"
        )
