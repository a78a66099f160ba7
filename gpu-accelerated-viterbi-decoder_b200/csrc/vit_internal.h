// vit_internal.h -- shared between the translation units of libvitb200.so; nothing here is exported.
#pragma once
// records `msg` as the calling thread's vit_last_error() text and returns `code`
int vit_set_error(int code, const char* msg);
