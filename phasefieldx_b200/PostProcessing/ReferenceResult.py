"""`AllResults`: loader of a results folder, API of reference
`src/phasefieldx/PostProcessing/ReferenceResult.py:16-258` (the reference tests read results
through it, e.g. `S.reaction_files["top.reaction"]["Ry"]`)."""
import os

import pandas as pd

from phasefieldx_b200.files import get_filenames_in_folder, get_foldernames_in_folder

_KINDS = {".conv": "convergence_files", ".energy": "energy_files", ".reaction": "reaction_files", ".dof": "dof_files"}


class AllResults:
    def __init__(self, folder_path):
        self.folder_path = folder_path
        self.input = {}
        self.convergence_files, self.energy_files, self.reaction_files, self.dof_files = {}, {}, {}, {}
        self.label, self.color = "label", "k"
        self.file_names = get_filenames_in_folder(folder_path)
        for name in self.file_names:
            path = os.path.join(folder_path, name)
            ext = os.path.splitext(name)[1]
            if ext in _KINDS:
                getattr(self, _KINDS[ext])[name] = pd.read_csv(path, sep="\t")
            elif ext == ".input":
                df = pd.read_csv(path, sep="\t", header=None, comment="/")
                self.input[name] = dict(zip(df[0], df[1]))
        self.folder_names = get_foldernames_in_folder(folder_path)
        vtu = os.path.join(folder_path, "paraview-solutions_vtu")
        if os.path.isdir(vtu):
            self.paraview_filenames = get_filenames_in_folder(vtu)

    def set_label(self, label):
        self.label = label

    def set_color(self, color):
        self.color = color

    def __str__(self):
        return f"AllResults({self.folder_path!r}: {len(self.file_names)} files)"
