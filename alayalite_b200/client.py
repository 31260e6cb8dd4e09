"""Client: a name -> Index registry with optional on-disk persistence.

Covers the index half of the reference's Client (/root/reference/python/src/alayalite/client.py:31-294):
create_index / get_index / get_or_create_index / list_indices / delete_index / save_index / reset and
reload-on-construction. Collections (documents + metadata in pandas) sit above the hot path and are
out of scope; the reference's Collection works unchanged on top of Index.
"""

from __future__ import annotations

import json
import os
import shutil

from .index import Index
from .schema import IndexParams, is_index_url

__all__ = ["Client"]


class Client:
    def __init__(self, url=None):
        self._indices = {}
        self._url = None
        if url is None:
            return
        self._url = os.path.abspath(url)
        os.makedirs(self._url, exist_ok=True)
        for name in sorted(os.listdir(self._url)):
            folder = os.path.join(self._url, name)
            if os.path.isdir(folder) and is_index_url(folder):
                self._indices[name] = Index.load(self._url, name)

    def list_indices(self):
        return list(self._indices)

    def get_index(self, name: str = "default"):
        return self._indices.get(name)

    def create_index(self, name: str = "default", **kwargs) -> Index:
        if name in self._indices:
            raise RuntimeError(f"A collection or index with name '{name}' already exists")
        self._indices[name] = Index(name, IndexParams.from_kwargs(**kwargs))
        return self._indices[name]

    def get_or_create_index(self, name: str, **kwargs) -> Index:
        return self._indices[name] if name in self._indices else self.create_index(name, **kwargs)

    def _need_url(self, message="Client is not initialized with a url for disk operations"):
        if self._url is None:
            raise RuntimeError(message)

    def delete_index(self, index_name: str, delete_on_disk: bool = False):
        if index_name not in self._indices:
            raise RuntimeError(f"Index '{index_name}' does not exist")
        del self._indices[index_name]
        if delete_on_disk:
            self._need_url()
            shutil.rmtree(os.path.join(self._url, index_name), ignore_errors=True)

    def save_index(self, index_name: str):
        self._need_url("Client is not initialized with a url")
        if index_name not in self._indices:
            raise RuntimeError(f"Index '{index_name}' does not exist")
        folder = os.path.join(self._url, index_name)
        schema = self._indices[index_name].save(folder)
        with open(os.path.join(folder, "schema.json"), "w", encoding="utf-8") as f:
            json.dump(schema, f, indent=4)

    def reset(self, delete_on_disk: bool = False):
        if delete_on_disk:
            self._need_url()
            for name in self._indices:
                shutil.rmtree(os.path.join(self._url, name), ignore_errors=True)
        self._indices = {}
