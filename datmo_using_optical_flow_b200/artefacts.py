"""The files the reference's driver loop writes per frame pair, in the reference's formats.

Restates the data-bearing half of /root/reference/Optical_flow/saving_utils.py (the matplotlib
PNGs are out of scope):
  bev_frame_{i}.npy                         saving_utils.py:65-66   save_bev
  velocity_x_frame_{i}.npy / velocity_y_..  saving_utils.py:69-71   save_velocity_grid
  dbscan_labels_frame_{i}.npy / _indices_.. saving_utils.py:106-108 save_dbscan_results
  ekf_tracks_frame_{i}.yaml                 saving_utils.py:119-125 save_ekf_tracks  ({track id: state list})
  <tracks>.csv                              saving_utils.py:80-103  save_all_velocities_to_csv
  <filtered velocities>.csv                 saving_utils.py:17-46   save_all_filtered_velocities_to_csv
Same file names, array dtypes, CSV headers, row order and number formatting (csv.writer's str() of
numpy scalars), so downstream scripts that read the reference's output directory keep working.
"""
from __future__ import annotations

import csv
import os

import numpy as np
import yaml

TRACK_CSV_HEADER = ["Frame Index", "Track ID", "Linear Velocity", "X Velocity", "Y Velocity", "Angular Velocity"]
CELL_CSV_HEADER = ["Frame Index", "Point Index", "Filtered X Velocity", "Filtered Y Velocity", "Magnitude",
                   "Angular Velocity"]


def save_bev(output_dir, bev, frame_index):
    np.save(os.path.join(output_dir, f"bev_frame_{frame_index}.npy"), bev)


def save_velocity_grid(output_dir, vx, vy, frame_index):
    np.save(os.path.join(output_dir, f"velocity_x_frame_{frame_index}.npy"), vx)
    np.save(os.path.join(output_dir, f"velocity_y_frame_{frame_index}.npy"), vy)


def save_dbscan_results(output_dir, labels, valid_indices, frame_index):
    np.save(os.path.join(output_dir, f"dbscan_labels_frame_{frame_index}.npy"), labels)
    np.save(os.path.join(output_dir, f"dbscan_indices_frame_{frame_index}.npy"), valid_indices)


def _states(tracks):
    """{id: state (4,)} from a TrackManager, a {id: Track} dict or a {id: array} dict."""
    tracks = getattr(tracks, "tracks", tracks)
    return {tid: np.asarray(getattr(t, "state", t), dtype=np.float64) for tid, t in tracks.items()}


def save_ekf_tracks(output_dir, tracks, frame_index):
    data = {tid: s.tolist() for tid, s in _states(tracks).items()}
    with open(os.path.join(output_dir, f"ekf_tracks_frame_{frame_index}.yaml"), "w") as fh:
        yaml.dump(data, fh)


def save_all_velocities_to_csv(tracks, frame_index, csv_file):
    """One row per track: |(state[2], state[3])|, state[2], state[3], state[1] (the reference labels
    state[1] "angular velocity"; reproduced, not fixed)."""
    exists = os.path.exists(csv_file)
    with open(csv_file, mode="a", newline="") as fh:
        w = csv.writer(fh)
        if not exists:
            w.writerow(TRACK_CSV_HEADER)
        for tid, s in _states(tracks).items():
            w.writerow([frame_index, tid, np.linalg.norm(s[2:4]), s[2], s[3], s[1]])


def save_all_filtered_velocities_to_csv(vx_filtered, vy_filtered, magnitude, angular_velocity, frame_index, csv_file):
    """One row per cell with a non-zero filtered velocity, in row-major order."""
    exists = os.path.exists(csv_file)
    with open(csv_file, mode="a", newline="") as fh:
        w = csv.writer(fh)
        if not exists:
            w.writerow(CELL_CSV_HEADER)
        ii, jj = np.nonzero((vx_filtered != 0) | (vy_filtered != 0))
        for idx, (i, j) in enumerate(zip(ii, jj)):
            w.writerow([frame_index, idx, vx_filtered[i, j], vy_filtered[i, j], magnitude[i, j], angular_velocity[i, j]])
