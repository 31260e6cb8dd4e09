"""Client: a name -> Index / Collection registry with optional on-disk persistence.

Mirrors the reference's Client (/root/reference/python/src/alayalite/client.py:31-294): create / get /
get_or_create / list / delete / save for indices and collections, reset, and reload-on-construction.
"""

from __future__ import annotations

import json
import os
import shutil

from .collection import Collection
from .index import Index
from .schema import IndexParams, is_collection_url, is_index_url

__all__ = ["Client"]


class Client:
    def __init__(self, url=None):
        self._indices = {}
        self._collections = {}
        self._url = None
        if url is None:
            return
        self._url = os.path.abspath(url)
        os.makedirs(self._url, exist_ok=True)
        for name in sorted(os.listdir(self._url)):
            folder = os.path.join(self._url, name)
            if not os.path.isdir(folder):
                continue
            if is_collection_url(folder):
                self._collections[name] = Collection.load(self._url, name)
            elif is_index_url(folder):
                self._indices[name] = Index.load(self._url, name)

    def _free_name(self, name):
        if name in self._indices or name in self._collections:
            raise RuntimeError(f"A collection or index with name '{name}' already exists")

    def _save(self, registry, kind, name):
        self._need_url("Client is not initialized with a url")
        if name not in registry:
            raise RuntimeError(f"{kind} '{name}' does not exist")
        folder = os.path.join(self._url, name)
        schema = registry[name].save(folder)
        with open(os.path.join(folder, "schema.json"), "w", encoding="utf-8") as f:
            json.dump(schema, f, indent=4)

    def _delete(self, registry, kind, name, delete_on_disk):
        if name not in registry:
            raise RuntimeError(f"{kind} '{name}' does not exist")
        del registry[name]
        if delete_on_disk:
            self._need_url()
            shutil.rmtree(os.path.join(self._url, name), ignore_errors=True)

    # -- collections ------------------------------------------------------------------------------
    def list_collections(self):
        return list(self._collections)

    def get_collection(self, name: str = "default"):
        return self._collections.get(name)

    def create_collection(self, name: str = "default", **kwargs) -> Collection:
        self._free_name(name)
        self._collections[name] = Collection(name, IndexParams.from_kwargs(**kwargs))
        return self._collections[name]

    def get_or_create_collection(self, name: str, **kwargs) -> Collection:
        return self._collections[name] if name in self._collections else self.create_collection(name, **kwargs)

    def delete_collection(self, collection_name: str, delete_on_disk: bool = False):
        self._delete(self._collections, "Collection", collection_name, delete_on_disk)

    def save_collection(self, collection_name: str):
        self._save(self._collections, "Collection", collection_name)

    # -- indices ----------------------------------------------------------------------------------

    def list_indices(self):
        return list(self._indices)

    def get_index(self, name: str = "default"):
        return self._indices.get(name)

    def create_index(self, name: str = "default", **kwargs) -> Index:
        self._free_name(name)
        self._indices[name] = Index(name, IndexParams.from_kwargs(**kwargs))
        return self._indices[name]

    def get_or_create_index(self, name: str, **kwargs) -> Index:
        return self._indices[name] if name in self._indices else self.create_index(name, **kwargs)

    def _need_url(self, message="Client is not initialized with a url for disk operations"):
        if self._url is None:
            raise RuntimeError(message)

    def delete_index(self, index_name: str, delete_on_disk: bool = False):
        self._delete(self._indices, "Index", index_name, delete_on_disk)

    def save_index(self, index_name: str):
        self._save(self._indices, "Index", index_name)

    def reset(self, delete_on_disk: bool = False):
        if delete_on_disk:
            self._need_url()
            for name in list(self._indices) + list(self._collections):
                shutil.rmtree(os.path.join(self._url, name), ignore_errors=True)
        self._indices = {}
        self._collections = {}
