// vit_internal.h -- shared between the translation units of libvitb200.so; nothing here is exported.
#pragma once
// records `msg` as the calling thread's vit_last_error() text and returns `code`
int vit_set_error(int code, const char* msg);

// Device memory that lives on ANOTHER GPU although this process addresses it (CUDA IPC mappings and peer pointers handed out by
// vit_comm_shared_alloc): cudaPointerGetAttributes reports an IPC mapping as local memory, so the communicator registers the
// ranges and the launcher asks here.  A decode whose output lies in such a range stages its stores (KParams::stage_out).
#include <stddef.h>
// owner_device: the CUDA ordinal (in this process) of the GPU that holds the memory, or -1 for "not a device of this
// process's numbering" (IPC mapping): the range is remote for every device but its owner.
void vit_note_remote_range(const void* base, size_t bytes, int owner_device);
void vit_forget_remote_range(const void* base);
bool vit_in_remote_range(const void* p, int device);
