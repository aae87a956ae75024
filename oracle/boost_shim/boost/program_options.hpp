// boost::program_options stand-in covering what reseq/main.cpp uses (oracle build only).
#ifndef RSQ_SHIM_BOOST_PROGRAM_OPTIONS_HPP
#define RSQ_SHIM_BOOST_PROGRAM_OPTIONS_HPP
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace boost { namespace program_options {

class error : public std::logic_error {
public:
	explicit error(const std::string &w) : std::logic_error(w) {}
};

// type-erased holder -----------------------------------------------------------------------------
class variable_value {
	struct holder_base { virtual ~holder_base() {} virtual const std::type_info &type() const = 0; };
	template<class T> struct holder : holder_base {
		T v; explicit holder(const T &x) : v(x) {}
		const std::type_info &type() const override { return typeid(T); }
	};
	std::shared_ptr<holder_base> h_;
	bool defaulted_ = false;
public:
	variable_value() {}
	template<class T> variable_value(const T &v, bool defaulted) : h_(new holder<T>(v)), defaulted_(defaulted) {}
	bool empty() const { return !h_; }
	bool defaulted() const { return defaulted_; }
	template<class T> const T &as() const {
		if(!h_ || h_->type() != typeid(T)){ throw error("boost::bad_any_cast: failed conversion using boost::any_cast"); }
		return static_cast<const holder<T> *>(h_.get())->v;
	}
};

class variables_map : public std::map<std::string, variable_value> {};

// value semantics -------------------------------------------------------------------------------
class value_semantic {
public:
	virtual ~value_semantic() {}
	virtual bool has_default() const = 0;
	virtual variable_value default_value_holder() const = 0;
	virtual variable_value parse(const std::string &name, const std::string &text) const = 0;
	virtual void notify(const variable_value &v) const = 0;
	virtual std::string default_text() const = 0;
};

namespace shim_detail {
	template<class T> inline bool lexical(const std::string &s, T &out){
		std::istringstream is(s); is >> out; return !is.fail() && (is.peek() == std::char_traits<char>::eof());
	}
	inline bool lexical(const std::string &s, std::string &out){ out = s; return true; }
	// boost::lexical_cast treats 8-bit integer targets as characters; main.cpp never uses them
}

template<class T> class typed_value : public value_semantic {
	T *store_to_;
	bool has_default_ = false;
	T default_;
public:
	explicit typed_value(T *store_to) : store_to_(store_to) {}
	typed_value *default_value(const T &v){ has_default_ = true; default_ = v; return this; }
	bool has_default() const override { return has_default_; }
	variable_value default_value_holder() const override { return variable_value(default_, true); }
	variable_value parse(const std::string &name, const std::string &text) const override {
		T v;
		if(!shim_detail::lexical(text, v)){ throw error("the argument ('" + text + "') for option '--" + name + "' is invalid"); }
		return variable_value(v, false);
	}
	void notify(const variable_value &v) const override { if(store_to_){ *store_to_ = v.as<T>(); } }
	std::string default_text() const override { std::ostringstream os; os << default_; return os.str(); }
};

template<class T> inline typed_value<T> *value(){ return new typed_value<T>(nullptr); }
template<class T> inline typed_value<T> *value(T *v){ return new typed_value<T>(v); }

// option descriptions ---------------------------------------------------------------------------
struct option_description {
	std::string long_name;
	char short_name = 0;
	std::string description;
	std::shared_ptr<const value_semantic> semantic; // null: flag without argument
};

class options_description;
class options_description_easy_init {
	options_description *owner_;
public:
	explicit options_description_easy_init(options_description *o) : owner_(o) {}
	options_description_easy_init &operator()(const char *name, const char *description);
	options_description_easy_init &operator()(const char *name, const value_semantic *s, const char *description);
};

class options_description {
	std::string caption_;
	std::vector<std::shared_ptr<option_description>> options_;
	std::vector<options_description> groups_;
	friend class options_description_easy_init;
public:
	explicit options_description(const std::string &caption = std::string()) : caption_(caption) {}
	options_description_easy_init add_options(){ return options_description_easy_init(this); }
	options_description &add(const options_description &d){ groups_.push_back(d); return *this; }

	void collect(std::vector<std::shared_ptr<option_description>> &out) const {
		for(const auto &o : options_){ out.push_back(o); }
		for(const auto &g : groups_){ g.collect(out); }
	}
	const option_description *find(const std::string &long_name, char short_name) const {
		std::vector<std::shared_ptr<option_description>> all; collect(all);
		for(const auto &o : all){
			if(short_name ? (o->short_name == short_name) : (o->long_name == long_name)){ return o.get(); }
		}
		if(!short_name){ // unambiguous prefix, as boost allows guessing
			const option_description *hit = nullptr; unsigned n = 0;
			for(const auto &o : all){ if(0 == o->long_name.compare(0, long_name.size(), long_name)){ hit = o.get(); ++n; } }
			if(1 == n){ return hit; }
		}
		return nullptr;
	}
	void print(std::ostream &os) const {
		if(!caption_.empty()){ os << caption_ << ":\n"; }
		for(const auto &o : options_){
			std::string left = "  ";
			if(o->short_name){ left += std::string("-") + o->short_name + " [ --" + o->long_name + " ]"; }
			else{ left += "--" + o->long_name; }
			if(o->semantic){ left += " arg"; if(o->semantic->has_default()){ left += " (=" + o->semantic->default_text() + ")"; } }
			if(left.size() < 38){ left.resize(38, ' '); } else { left += ' '; }
			os << left << o->description << '\n';
		}
		for(const auto &g : groups_){ os << '\n'; g.print(os); }
	}
};
inline std::ostream &operator<<(std::ostream &os, const options_description &d){ d.print(os); return os; }

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *description){
	return (*this)(name, nullptr, description);
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const value_semantic *s, const char *description){
	auto o = std::make_shared<option_description>();
	std::string n(name);
	auto comma = n.find(',');
	if(std::string::npos != comma){ o->long_name = n.substr(0, comma); o->short_name = n[comma+1]; }
	else{ o->long_name = n; }
	o->description = description;
	o->semantic.reset(s);
	owner_->options_.push_back(o);
	return *this;
}

// parsing --------------------------------------------------------------------------------------
struct option {
	std::string string_key;
	int position_key = -1;
	std::vector<std::string> value;
	std::vector<std::string> original_tokens;
	bool unregistered = false;
};
struct parsed_options {
	std::vector<option> options;
	const options_description *description = nullptr;
};

enum collect_unrecognized_mode { include_positional, exclude_positional };

class command_line_parser {
	std::vector<std::string> args_;
	const options_description *desc_ = nullptr;
	bool allow_unregistered_ = false;
public:
	command_line_parser(int argc, const char *const argv[]){ for(int i=1; i<argc; ++i){ args_.push_back(argv[i]); } }
	explicit command_line_parser(const std::vector<std::string> &args) : args_(args) {}
	command_line_parser &options(const options_description &d){ desc_ = &d; return *this; }
	command_line_parser &allow_unregistered(){ allow_unregistered_ = true; return *this; }

	parsed_options run(){
		parsed_options result; result.description = desc_;
		int position = 0;
		for(std::size_t i=0; i<args_.size(); ++i){
			const std::string &tok = args_[i];
			option opt;
			opt.original_tokens.push_back(tok);
			const option_description *d = nullptr;
			std::string attached; bool has_attached = false;

			if(tok.size() > 2 && '-' == tok[0] && '-' == tok[1]){
				std::string name = tok.substr(2);
				auto eq = name.find('=');
				if(std::string::npos != eq){ attached = name.substr(eq+1); has_attached = true; name = name.substr(0, eq); }
				d = desc_->find(name, 0);
				opt.string_key = d ? d->long_name : name;
			}
			else if(tok.size() > 1 && '-' == tok[0] && '-' != tok[1]){
				d = desc_->find(std::string(), tok[1]);
				opt.string_key = d ? d->long_name : std::string(1, tok[1]);
				if(tok.size() > 2){ attached = tok.substr(2); has_attached = true; }
				if(!d){ opt.string_key = "-" + tok.substr(1, 1); }
			}
			else{
				// positional
				opt.position_key = position++;
				opt.value.push_back(tok);
				opt.unregistered = true; // no positional_options_description is ever given
				if(!allow_unregistered_){ throw error("too many positional options have been specified on the command line"); }
				result.options.push_back(opt);
				continue;
			}

			if(!d){
				if(!allow_unregistered_){ throw error("unrecognised option '" + tok + "'"); }
				opt.unregistered = true;
				if(has_attached){ opt.value.push_back(attached); }
				result.options.push_back(opt);
				continue;
			}

			if(d->semantic){
				if(has_attached){ opt.value.push_back(attached); }
				else{
					if(i+1 >= args_.size()){ throw error("the required argument for option '--" + d->long_name + "' is missing"); }
					opt.value.push_back(args_[++i]);
					opt.original_tokens.push_back(args_[i]);
				}
			}
			else if(has_attached && tok[1] != '-'){
				// bundled short flags are not used by reseq; treat the remainder as an error
				throw error("option '--" + d->long_name + "' does not take any arguments");
			}
			result.options.push_back(opt);
		}
		return result;
	}
};

inline std::vector<std::string> collect_unrecognized(const std::vector<option> &options, collect_unrecognized_mode mode){
	std::vector<std::string> out;
	for(const auto &o : options){
		if(o.unregistered || (include_positional == mode && -1 != o.position_key)){
			if(-1 == o.position_key || include_positional == mode){
				out.insert(out.end(), o.original_tokens.begin(), o.original_tokens.end());
			}
		}
	}
	return out;
}

inline void store(const parsed_options &parsed, variables_map &vm){
	for(const auto &o : parsed.options){
		if(o.unregistered){ continue; }
		const option_description *d = parsed.description->find(o.string_key, 0);
		if(!d){ continue; }
		if(vm.count(d->long_name) && !vm[d->long_name].defaulted()){
			throw error("option '--" + d->long_name + "' cannot be specified more than once");
		}
		if(d->semantic){ vm[d->long_name] = d->semantic->parse(d->long_name, o.value.at(0)); }
		else{ vm[d->long_name] = variable_value(std::string(), false); }
	}
	// defaults + remember semantics for notify
	std::vector<std::shared_ptr<option_description>> all; parsed.description->collect(all);
	for(const auto &d : all){
		if(d->semantic && d->semantic->has_default() && !vm.count(d->long_name)){
			vm[d->long_name] = d->semantic->default_value_holder();
		}
		if(d->semantic && vm.count(d->long_name)){ d->semantic->notify(vm[d->long_name]); }
	}
}
inline void notify(variables_map &){ /* store_to targets are written in store() */ }

}}
#endif
