{
  # node-gyp build description for the N-API shim (NOT built in this repository's image: no Node).
  "targets": [{
    "target_name": "jsmpeg_b200",
    "sources": ["jsmpeg_b200_napi.c"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../jsmpeg_b200", "-ljsmpeg_b200",
                  "-Wl,-rpath,<(module_root_dir)/../jsmpeg_b200"]
  }]
}
