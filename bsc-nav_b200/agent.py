"""Retrieval slice of the agent: the calls that enter the spatial memory.

GESObjectNavRobot.working_memory_retrival / weighted_cluster_centers / _grid2loc / _loc2grid (BSCAgent.py:454-497,
587-664; twin in objnav_benchmark.py:440-495) and LocalizeAgent (agent_localize.py:11-37).  LLM prompt expansion,
Motion execution and the LLM are out-of-scope collaborators: `prompt_expander` is injectable.  The CLIP view check of
check_around (BSCAgent.py:700-748) is `ViewChecker`: image tower + scoring on the device, text tower supplied by the caller.
"""
from __future__ import annotations

import os

import numpy as np


class GESObjectNavRobot:
    def __init__(self, memory, habitat_benchmark_env=None, prompt_expander=None):
        self.memory = memory
        self.benchmark_env = habitat_benchmark_env          # may be None (agent_localize.py:12-13)
        self.prompt_expander = prompt_expander              # stands in for imagenary_helper(client, text) (LLMAgent.py:70)

    # -- BSCAgent.py:454-477
    def _agent_position(self):
        env = self.benchmark_env
        sim = getattr(env, "sim", None) or getattr(env, "sims", None)      # both spellings exist
        return sim.agents[0].get_state().position

    def _grid2loc(self, grid_id):
        row, col, _height = grid_id[0], grid_id[1], grid_id[2]
        initial_x, _initial_z, initial_y = self.memory.Env.original_state.position
        actual_y = initial_y + (row - self.memory.gs // 2) * self.memory.cs
        actual_x = initial_x + (col - self.memory.gs // 2) * self.memory.cs
        actual_z = self._agent_position()[1] + 0.2
        return np.array([actual_x, actual_z, actual_y])

    def _loc2grid(self, loc):
        x, z, y = loc
        initial_x, _initial_z, initial_y = self.memory.Env.original_state.position
        row = int((y - initial_y) / self.memory.cs + self.memory.gs // 2)
        col = int((x - initial_x) / self.memory.cs + self.memory.gs // 2)
        h = int(z / self.memory.cs) - self.memory.minh
        return np.array([row, col, h])

    # -- BSCAgent.py:479-497
    def weighted_cluster_centers(self, top_k_positions, top_k_similarity, eps=10, min_samples=5):
        """Same signature and return values as the reference (host arrays in and out); DBSCAN, the weighted centres and the
        ordering run in bscnav_cluster_centers on the device (csrc/cluster.cu) -- no scikit-learn on the product path."""
        import torch
        pos = np.asarray(top_k_positions).reshape(-1, 3)
        sim = np.asarray(top_k_similarity, dtype=np.float32).reshape(-1)
        n = len(sim)
        if n == 0:
            return np.zeros((0,)), np.zeros((0,), dtype=np.int64), []
        store = self.memory.store
        dev = store.device
        out = store.cluster_centers(torch.from_numpy(sim).to(dev)[None], torch.from_numpy(pos.astype(np.int32)).to(dev)[None],
                                    torch.tensor([n], dtype=torch.int32, device=dev), eps=eps, min_samples=min_samples,
                                    max_clusters=max(1, min(n, 256)))
        centers, _mean, sizes, labels, n_c = (t[0].cpu().numpy() for t in out)
        n_c = int(n_c)
        return centers[:n_c].copy(), labels.astype(np.int64), [int(v) for v in sizes[:n_c]]

    # -- BSCAgent.py:587-664
    def working_memory_retrival(self, prompts, vis_aug=False, text_aug=True, region_radius=np.inf, curr_grid=None):
        if curr_grid is None and region_radius != np.inf:
            curr_grid = self._loc2grid(self._agent_position())
        if isinstance(prompts, (str, list)):
            text = prompts if isinstance(prompts, str) else " ".join(prompts)
            if text_aug and self.prompt_expander is not None:
                text = self.prompt_expander(text)
            _best, top_pos, top_sim = self.memory.voxel_localized(text, region_radius=region_radius, curr_grid=curr_grid)
        else:
            _best, top_pos, top_sim = self.memory.voxel_localized(prompts, region_radius=region_radius, curr_grid=curr_grid)
        centers, _, _ = self.weighted_cluster_centers(top_pos, top_sim)
        if isinstance(prompts, str):
            name = prompts[:64]
            os.makedirs(self.memory.memory_save_path, exist_ok=True)
            np.save(self.memory.memory_save_path + f"/best_pos_topK_{name}.npy", np.array(top_pos))
            np.save(self.memory.memory_save_path + f"/best_pos_centers_{name}.npy", np.array(centers))
        return np.array([centers])


class LocalizeAgent(GESObjectNavRobot):
    """agent_localize.py:11-37."""

    def __init__(self, memory, habitat_benchmark_env=None, prompt_expander=None, save_dir="localize_results"):
        super().__init__(memory, habitat_benchmark_env, prompt_expander)
        self.save_dir = save_dir
        os.makedirs(self.save_dir, exist_ok=True)

    def LocalizeTextPrompt(self, text_prompt):
        text = self.prompt_expander(text_prompt) if self.prompt_expander else text_prompt
        _best, top_pos, top_sim = self.memory.voxel_localized(text)
        centers, _, _ = self.weighted_cluster_centers(top_pos, top_sim)
        np.save(f"{self.save_dir}/best_pos_topK_text_prompt.npy", np.array(top_pos))
        return centers

    def LocalizeImagePrompt(self, img):
        _best, top_pos, top_sim = self.memory.voxel_localized(img)
        centers, _, _ = self.weighted_cluster_centers(top_pos, top_sim)
        np.save(f"{self.save_dir}/best_pos_topK_img_input.npy", np.array(top_pos))
        return centers


class ViewChecker:
    """The scoring half of check_around (BSCAgent.py:700-748): after a full turn the agent encodes every view with CLIP
    ViT-H/14, scores them against the goal (text or image embedding), takes the soft-max arg-max and turns back to that view.

    `clip` is a ViTEngine built from vit_config.OPENCLIP_H14 weights (image tower, `encode_image`).  Views and image goals are
    pre-processed tensors (N, 3, 224, 224) -- open_clip's PIL transform stays with the caller, as does the text tower: pass
    `encode_text` (str -> (1, E) tensor) for text goals."""

    def __init__(self, clip, turn_left: float, encode_text=None):
        self.clip, self.turn_left, self.encode_text = clip, turn_left, encode_text

    def goal_features(self, prompt):
        if isinstance(prompt, str):
            if self.encode_text is None:
                raise ValueError("text goal needs an `encode_text` collaborator (open_clip's text tower)")
            return self.encode_text(prompt).reshape(-1)
        return self.clip.encode_image(prompt.reshape(1, *prompt.shape[-3:]))[0]

    def best_view(self, views, prompt):
        """-> (index of the matching view, soft-max probabilities (N,) on the device)."""
        import torch
        from .vit import score_views
        if isinstance(prompt, str):
            feats, goal = self.clip.encode_image(views), self.goal_features(prompt)
        else:
            # an image goal rides in the views' batch: one pass of the tower instead of two (a lone image is launch-bound)
            goal_img = prompt.reshape(1, *prompt.shape[-3:]).to(views.device, views.dtype)
            both = self.clip.encode_image(torch.cat([views, goal_img], dim=0))
            feats, goal = both[:-1], both[-1]
        prob, best = score_views(feats, goal)
        return int(best), prob

    def turn_actions(self, max_idx: int):
        """BSCAgent.py:735-745: the turn that faces the best view again."""
        num_turns = int(360 / self.turn_left)
        if max_idx >= num_turns:
            max_idx -= num_turns
        if max_idx * self.turn_left <= 180:
            return ["turn_left"] * max_idx
        return ["turn_right"] * (num_turns - max_idx)


def match_long_memory_label(text_features, label_features) -> int:
    """Long-term label matching (BSCAgent.py:541-557): the stored label whose CLIP text embedding is closest (cosine) to the
    prompt's; embeddings from a text-tower ViTEngine (`encode_text`), arg-max on the device."""
    from .vit import score_views
    _prob, best = score_views(label_features, text_features.reshape(-1))
    return int(best)
