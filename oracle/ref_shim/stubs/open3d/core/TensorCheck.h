#pragma once
#define AssertTensorShape(...) ref_shim_noop()
#define AssertTensorDtype(...) ref_shim_noop()
#define AssertTensorDtypes(...) ref_shim_noop()
#define AssertTensorDevice(...) ref_shim_noop()
namespace open3d { namespace core { inline void ref_shim_noop() {} } }
