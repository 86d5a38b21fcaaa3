"""
Result-file utilities with the API and on-disk formats of reference `src/phasefieldx/files.py`
(`prepare_simulation` :324-363 wipes and recreates the results folder; `append_results_to_file`
:409-439 appends one TSV row, writing the header when the file is empty; values are `str()` of
Python objects).  Only paths inside `path` are ever deleted.
"""
import os
import shutil
from pathlib import Path
from typing import List

_STALE_SUFFIXES = (".log", ".conv", ".h5", ".xdmf", ".energy", ".reaction", ".dof")


def get_filenames_in_folder(folder_path: str) -> List[str]:
    return [p.name for p in Path(folder_path).iterdir() if p.is_file()]


def get_foldernames_in_folder(folder_path: str) -> List[str]:
    return [p.name for p in Path(folder_path).iterdir() if p.is_dir()]


def find_files_with_extension(folder_path: str, extension: str) -> List[str]:
    hits = []
    for root, _dirs, names in os.walk(folder_path):
        hits.extend(os.path.join(root, n) for n in names if n.endswith(extension))
    return sorted(hits)


def read_last_line(file_path):
    rows = Path(file_path).read_text().splitlines()
    return rows[-1].strip() if rows else None


def read_last_lines(file_path, num_lines):
    rows = Path(file_path).read_text().splitlines()
    return [r.strip() for r in rows[-num_lines:]]


def replace_text(original, target, list_replace):
    """Copy `original` to `target` applying (pattern, replacement) pairs; True on success."""
    try:
        text = Path(original).read_text()
        for pattern, replacement in list_replace:
            text = text.replace(pattern, replacement)
        Path(target).write_text(text)
        return True
    except Exception as exc:  # same contract as the reference: report and return False
        print(f"An error occurred while replacing text: {exc}")
        return False


def delete_specific_files(folder_path, extensions):
    try:
        for p in Path(folder_path).iterdir():
            if p.is_file() and p.suffix in extensions:
                p.unlink()
                print(f"Deleted: {p}")
        print("All specified files deleted successfully.")
    except Exception as exc:
        print(f"An error occurred: {exc}")


def delete_folders_with_contents(folder_path, folder_names):
    try:
        for name in folder_names:
            target = os.path.join(folder_path, name)
            if os.path.isdir(target):
                shutil.rmtree(target)
                print(f"Deleted folder and its contents: {target}")
        print("All specified folders and their contents deleted successfully.")
    except Exception as exc:
        print(f"An error occurred: {exc}")


def prepare_simulation(path, results_folder_name="results"):
    """Remove stale result files/folders under `path` and create an empty results folder."""
    delete_specific_files(path, list(_STALE_SUFFIXES))
    delete_folders_with_contents(path, ["paraview-solutions"])
    results_path = os.path.join(path, results_folder_name)
    delete_folders_with_contents(path, [results_path])
    os.makedirs(results_path)


def delete_folder(folder_path):
    try:
        shutil.rmtree(folder_path)
        print(f"Folder '{folder_path}' and its contents deleted successfully.")
    except Exception as exc:
        print(f"Error: {exc}")


def append_results_to_file(output_file_path, header, step, *data):
    """Append `step<TAB>data...` (str() of each value); the header goes first into an empty file."""
    with open(output_file_path, "a") as fh:
        if fh.tell() == 0 and os.path.getsize(output_file_path) == 0:
            fh.write(f"{header}\n")
        fh.write("\t".join([str(step)] + [str(d) for d in data]) + "\n")
